"""Shared helpers of the GPU parity tests: build a drop-in model on the GPU from a golden case."""
import numpy as np
import scipy.sparse as sp
import torch

from oracle import cf_oracle as O
from oracle import inputs, replay


def make_model(model_key, case, hp, inject=None, device='cuda'):
    import sslrec_b200
    from sslrec_b200.config import default_config, load_config
    from sslrec_b200.data_handler import DataHandlerGeneralCF
    name = model_key.split('_')[0]
    cfg = default_config(name, **hp)
    cfg['model']['embedding_size'] = case['dim']
    cfg['train']['batch_size'] = case['batch']
    if name == 'ncl':
        cfg['train']['loss'] = 'pairwise_with_epoch_flag'
    load_config(base=cfg, device=device)
    trn = sp.coo_matrix((np.ones(len(case['rows']), dtype=np.float32), (case['rows'], case['cols'])),
                        shape=(case['n_user'], case['n_item']))
    dh = DataHandlerGeneralCF(trn)
    dh.load_data()
    import importlib
    mod = importlib.import_module('sslrec_b200.general_cf.' + name)
    cls = [getattr(mod, a) for a in dir(mod) if a.lower() == name][0]
    model = cls(dh)
    model._inject = inject
    model = model.to(device)
    return model, dh


def gpu_injection(model_key, case, hp, adj, dr, device='cuda'):
    """replay.draws (oracle Adj order == CSR order) -> the model's ``_inject`` dict on the GPU."""
    name = model_key.split('_')[0]
    inj = {}
    u8 = lambda m: None if m is None else torch.from_numpy(np.asarray(m).astype(np.uint8)).to(device)
    if name == 'lightgcn':
        inj['edge_masks'] = [u8(dr['edge_keep'])] + [None] * 3
    elif name == 'simgcl':
        inj['noise_u'] = [[u.to(device).contiguous() for u in view] for view in dr['uniforms']]
    elif name == 'sgl':
        inj['edge_masks'] = [u8(m) for m in dr['edge_keeps']] + [None, None]
        inj['node_masks'] = [None if m is None else m.to(torch.uint8).to(device) for m in dr['node_keeps']]
    elif name == 'hccf':
        inj['edge_masks_per_layer'] = [u8(m) for m in dr['edge_keeps']]
        inj['hyper_keeps'] = [(ku.to(device), ki.to(device)) for ku, ki in dr['hyper_keeps']]
    return inj


def close(a, b, rtol, atol, what):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), f'{what}: {bad.sum()} / {bad.size} off, max err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)} (ref {b.flat[err.argmax()]:.4e})'
