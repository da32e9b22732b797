"""Trial-/fold-/cell-line-parallel hyper-parameter sweep over the GPUs of a node (BASELINE config 4).

The reference runs `Kfold_CV_Multimodal()(...)` serially: for each of 7 cell lines x {enhancers, promoters} = 14 data
sets, 3 folds, each a 3-trial study plus a final fit.  The 42 (data set, fold) jobs are independent, so the sweep
is "replicas only": one worker process per GPU, jobs handed out dynamically from a queue, NO collective; the shared
state is the append-only study log (hpo.JsonlStorage) and one JSON result line per job.

    python -m embrace_b200.sweep --gpus 8 --datasets 14 --rows 4096 --epochs 3 --out sweep_out

prints one JSON line: jobs, trials, wall seconds, trials/hour (synthetic data of the reference's shapes)."""
import argparse
import json
import os
import sys
import time

import numpy as np

CELL_F = {'A549': 48, 'GM12878': 152, 'H1': 58, 'HEK293': 196, 'HEPG2': 562, 'K562': 429, 'MCF7': 117}   # raw feature counts
SWEEP_TASKS = ['active_E_vs_inactive_E', 'active_P_vs_inactive_P']


def synthetic_dataset(cell_line, task, rows, seed):
    """(features [rows, F] in [0,1), base codes [rows, 256], labels) with a planted signal: logistic on 4 features and a
    6-mer motif in 70 % of the positives."""
    rs = np.random.RandomState(seed)
    F = CELL_F[cell_line]
    x = rs.random_sample((rows, F)).astype(np.float32)
    codes = rs.randint(0, 4, size=(rows, 256)).astype(np.uint8)
    z = 3.0 * (x[:, :4].sum(1) - 2.0) - 2.3
    y = (rs.random_sample(rows) < 1 / (1 + np.exp(-z))).astype(np.int64)
    motif = np.array([0, 2, 3, 3, 1, 0], dtype=np.uint8)
    for i in np.nonzero(y)[0]:
        if rs.random_sample() < 0.7:
            p = rs.randint(0, 250)
            codes[i, p:p + 6] = motif
    return x, codes, y


def make_jobs(n_datasets, n_folds):
    ds = [(c, t) for c in CELL_F for t in SWEEP_TASKS][:n_datasets]
    return [dict(cell_line=c, task=t, fold=f + 1) for (c, t) in ds for f in range(n_folds)]


def run_job(job, args, device):
    """One (data set, fold): a 3-trial study on train/validation, then the final fit on train+validation."""
    from .BIOINF_tesi.models import EmbraceNetMultimodal
    from .BIOINF_tesi.models.utils.training_models_multimodal import Kfold_CV_Multimodal, ArrayPipeline
    seed = 1000 + 17 * (list(CELL_F).index(job['cell_line']) * 2 + SWEEP_TASKS.index(job['task']))
    x, codes, y = synthetic_dataset(job['cell_line'], job['task'], args.rows, seed)
    cv = Kfold_CV_Multimodal()
    t0 = time.time()
    cwd = os.getcwd()
    os.chdir(args.out)                            # trial / fold checkpoints are written relative to the CWD, as in the reference
    try:
        scores = cv(ArrayPipeline(x, codes, y), job['cell_line'], device, task=job['task'], model=EmbraceNetMultimodal,
                    n_folds=args.folds, num_epochs=args.epochs, batch_size=args.batch, n_trials=args.trials, sampler=args.sampler,
                    study_name=f"{job['cell_line']}_{job['task']}_EmbraceNetMultimodal", storage='sweep_studies.db',
                    sampler_seed=seed + job['fold'], folds=[job['fold']], test_model_path=None)
    finally:
        os.chdir(cwd)
    return dict(job, seconds=time.time() - t0, final_test_AUPRC=float(scores['final_test_AUPRC_scores'][-1]), trials=args.trials)


def worker(rank, args, jobs, results):
    import torch
    torch.cuda.set_device(rank)
    if not args.verbose:
        sys.stdout = open(os.devnull, 'w')
    while True:
        try:
            job = jobs.get_nowait()
        except Exception:
            break
        r = run_job(job, args, f'cuda:{rank}')
        r['gpu'] = rank
        results.put(r)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--datasets', type=int, default=14)
    ap.add_argument('--folds', type=int, default=3)
    ap.add_argument('--trials', type=int, default=3)
    ap.add_argument('--rows', type=int, default=4096, help='synthetic rows per data set')
    ap.add_argument('--epochs', type=int, default=3)
    ap.add_argument('--batch', type=int, default=100)
    ap.add_argument('--sampler', default='TPE')
    ap.add_argument('--out', default='sweep_out')
    ap.add_argument('--verbose', type=int, default=0)
    args = ap.parse_args(argv)
    import torch
    import torch.multiprocessing as mp
    os.makedirs(args.out, exist_ok=True)
    jobs_list = make_jobs(args.datasets, args.folds)
    t0 = time.time()
    out = []
    if args.gpus == 1:
        old = sys.stdout
        if not args.verbose:
            sys.stdout = open(os.devnull, 'w')
        try:
            for job in jobs_list:
                out.append(dict(run_job(job, args, 'cuda:0'), gpu=0))
        finally:
            sys.stdout = old
    else:
        ctx = mp.get_context('spawn')
        jobs, results = ctx.Queue(), ctx.Queue()
        for j in jobs_list:
            jobs.put(j)
        procs = [ctx.Process(target=worker, args=(r, args, jobs, results)) for r in range(args.gpus)]
        for p in procs:
            p.start()
        for _ in jobs_list:
            out.append(results.get())
        for p in procs:
            p.join()
    wall = time.time() - t0
    n_trials = sum(r['trials'] for r in out)
    line = {'metric': 'embracenet_sweep_trials_per_hour', 'value': 3600.0 * n_trials / wall, 'unit': 'trials/h', 'n_gpus': args.gpus,
            'jobs': len(out), 'trials': n_trials, 'final_fits': len(out), 'wall_s': wall, 'scaling': 'weak (replicas only, no collective)',
            'config': {'workload': f'{args.datasets} synthetic data sets x {args.folds} folds x ({args.trials} trials + final fit), '
                                   f'{args.rows} rows, {args.epochs} epochs, batch {args.batch} (BASELINE configs[3])'},
            'mean_final_test_AUPRC': float(np.mean([r['final_test_AUPRC'] for r in out])),
            'job_seconds': {'mean': float(np.mean([r['seconds'] for r in out])), 'max': float(np.max([r['seconds'] for r in out]))}}
    print(json.dumps(line), flush=True)
    return line


if __name__ == '__main__':
    main()
