"""Drive the UNMODIFIED reference (oracle/_ref, vendored by oracle/vendor_ref.py) on CPU for a graph given as arrays:
the harness of SURVEY.md 8(c) / BASELINE.md section 2.  TEST / BASELINE INFRASTRUCTURE ONLY -- used by
``bench.py --impl reference`` and the ``cpu_baseline`` leg; never by the product package.

A scratch CWD holds symlinks to the reference's config/ data_utils/ models/ trainer/ and a
datasets/general_cf/sparse_gowalla directory with the given matrix pickled the way the reference loads it
(data_handler_general_cf.py:12-35); ``sys.argv`` is set before ``config.configurator`` is imported (it parses at import,
configurator.py:57); hyper-parameters are overridden in ``configs['model']`` in place; ``Tensor.cuda`` is shimmed to
identity because aug_utils.py:130,147-154 hard-code ``.cuda()``.  One model per process (the reference's config is a
module-level singleton).  The step is trainer/trainer.py:63-68 verbatim in behaviour: zero_grad, cal_loss, loss.item(),
backward, Adam.step, float() of every loss term.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
import time

import numpy as np

from . import vendor_ref


class ReferenceTrainer:
    def __init__(self, model: str, rows, cols, n_user: int, n_item: int, hp: dict, batch: int = 4096, threads: int = None, csr: bool = False):
        import scipy.sparse as sp
        import torch
        ref = vendor_ref.vendor()
        self.scratch = d = tempfile.mkdtemp(prefix='sslrec_ref_')
        for sub in vendor_ref.DIRS:
            os.symlink(os.path.join(ref, sub), os.path.join(d, sub))
        dd = os.path.join(d, 'datasets', 'general_cf', 'sparse_gowalla')
        os.makedirs(dd)
        trn = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n_user, n_item))
        few = sp.coo_matrix((np.ones(8), (np.arange(8) % n_user, np.arange(8) % n_item)), shape=(n_user, n_item))
        for name, m in (('train_mat.pkl', trn), ('valid_mat.pkl', few), ('test_mat.pkl', few)):
            with open(os.path.join(dd, name), 'wb') as f:
                pickle.dump(m, f)
        self._cwd, self._argv = os.getcwd(), list(sys.argv)
        os.chdir(d)
        sys.path.insert(0, d)
        sys.argv = ['main.py', '--model', model, '--device', 'cpu']
        torch.Tensor.cuda = lambda self, *a, **k: self
        if threads:
            torch.set_num_threads(threads)
        from config.configurator import configs            # the reference's own module (from the scratch tree)
        known = set(configs['model'])
        configs['model'].update({k: v for k, v in hp.items() if k in known or k in ('embedding_size', 'layer_num')})
        configs['train']['batch_size'] = batch
        from data_utils.build_data_handler import build_data_handler
        from models.bulid_model import build_model
        from trainer.trainer import Trainer, init_seed
        init_seed()
        self.dh = build_data_handler()
        self.dh.load_data()
        self.model = build_model(self.dh).to(configs['device'])
        self.trainer = Trainer.__new__(Trainer)            # no Logger / Metric: only create_optimizer (trainer.py:45-49) is used
        self.trainer.data_handler = self.dh
        self.trainer.create_optimizer(self.model)
        self.configs = configs
        self.model.train()
        if csr:     # SURVEY.md 8(d) "tuned CPU": the one-line change adj.to_sparse_csr() (8.5x faster t.spmm on CPU); EdgeDrop needs COO
            if configs['model'].get('keep_rate', 1.0) != 1.0:
                raise ValueError('the CSR variant needs keep_rate = 1 (aug_utils.py:25-30 indexes the COO tensor)')
            self.model.adj = self.model.adj.coalesce().to_sparse_csr()

    def step(self, batch):
        """One iteration of trainer.py:63-68,71-72 on a batch of int64 tensors [ancs, poss, negs]."""
        opt = self.trainer.optimizer
        opt.zero_grad()
        batch_data = [x.long().to(self.configs['device']) for x in batch]
        if self.configs['model']['name'] == 'ncl' and len(batch_data) == 3:      # pairwise_with_epoch_flag datasets add the flag column
            batch_data.append(batch_data[0].new_zeros(batch_data[0].shape[0]) + (1 if not getattr(self, '_clustered', False) else 0))
            self._clustered = True
        loss, loss_dict = self.model.cal_loss(batch_data)
        v = loss.item()
        loss.backward()
        opt.step()
        for name in loss_dict:
            float(loss_dict[name])
        return v

    def close(self):
        os.chdir(self._cwd)
        sys.argv = self._argv


def time_steps(model, rows, cols, n_user, n_item, hp, batches, threads, budget_s, max_steps, warmup=1, csr=False):
    """([seconds per step], threads used) of up to ``max_steps`` reference steps inside ``budget_s``.  ``threads`` may be a
    list of candidates: one untimed step runs under each (doubling as warm-up) and the fastest setting is kept --
    torch's sparse COO addmm stops scaling early, so "all cores" is not always the reference's best case."""
    import torch
    cands = list(threads) if isinstance(threads, (list, tuple)) else [threads]
    tr = ReferenceTrainer(model, rows, cols, n_user, n_item, hp, batch=len(batches[0][0]), threads=cands[0], csr=csr)
    tb = [tuple(torch.from_numpy(np.asarray(b[i])) for i in range(3)) for b in batches]
    t_start = time.perf_counter()
    best, use = None, cands[0]
    for c in cands:
        if c:
            torch.set_num_threads(c)
        t0 = time.perf_counter()
        tr.step(tb[0])
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, use = dt, c
    if use:
        torch.set_num_threads(use)
    for i in range(max(0, warmup - len(cands))):
        tr.step(tb[i % len(tb)])
    times = []
    for i in range(max_steps):
        t0 = time.perf_counter()
        tr.step(tb[(warmup + i) % len(tb)])
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    tr.close()
    return times, (use or torch.get_num_threads())
