"""Mirror of BIOINF_tesi/models/utils/__init__.py:1-12."""
from .utils import (accuracy, AUPRC, F1_precision_recall, EarlyStopping, size_out_convolution, weight_reset,
                    get_loss_weights_from_dataloader, get_loss_weights_from_labels, output_size_from_model_params,
                    get_single_model_params)
from .training_models import fit
from .training_models_multimodal import fit_multimodal, Param_Search_Multimodal, Kfold_CV_Multimodal

__all__ = ['accuracy', 'AUPRC', 'F1_precision_recall', 'EarlyStopping', 'fit', 'size_out_convolution', 'weight_reset',
           'get_loss_weights_from_dataloader', 'get_loss_weights_from_labels', 'output_size_from_model_params', 'fit_multimodal',
           'get_single_model_params', 'Param_Search_Multimodal', 'Kfold_CV_Multimodal']
