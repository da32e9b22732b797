"""Mirror of fit (BIOINF_tesi/models/utils/training_models.py:31-186): the single-modality loop (FFNN / CNN)."""
import os

import torch

from .training_models_multimodal import lift_optimizer, run_epochs


def fit(model, train_loader, test_loader, device, optimizer=None, num_epochs=100, patience=4, delta=0, verbose=True,
        checkpoint_path=None):
    if os.path.exists(checkpoint_path):
        checkpoint = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
        model.load_state_dict(checkpoint['model_state_dict'])
        return checkpoint['AUPRC_train_scores'], checkpoint['AUPRC_test_scores'], checkpoint['F1_precision_recall_test_scores']
    cfg = lift_optimizer(optimizer)
    model = model.double().to(device)
    seq = model.spec.kind == 'cnn'

    def single(loader):
        def it():
            for data, target in loader:
                yield (None, data, target) if seq else (data, None, target)
        return it
    scores = run_epochs(model, single(train_loader), single(test_loader), len(train_loader), len(test_loader), cfg, num_epochs,
                        patience, delta, verbose)
    if checkpoint_path:
        torch.save({'model_state_dict': model.state_dict(), 'AUPRC_train_scores': scores[0], 'AUPRC_test_scores': scores[1],
                    'F1_precision_recall_test_scores': scores[2]}, checkpoint_path)
    return scores
