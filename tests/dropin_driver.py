"""Subprocess body of tests/test_dropin_reference.py: the reference's own main.py sequence (main.py:9-28) inside a scratch
SSLRec tree whose seven in-scope model modules are the one-line shims of INTEGRATION.md section 2.  Everything else --
config/configurator.py and the UNCHANGED config/modelconf/*.yml, data_utils (build_data_handler, DataHandlerGeneralCF,
datasets), models/bulid_model.py, trainer (build_trainer, Trainer.train_epoch, Metric.eval, Logger) -- is the reference's
code from oracle/_ref.  Prints one JSON line.  TEST INFRASTRUCTURE."""
import json
import os
import sys


def main():
    scratch, repo, model_name = sys.argv[1], sys.argv[2], sys.argv[3]
    os.chdir(scratch)
    sys.path.insert(0, scratch)
    sys.path.insert(1, repo)
    sys.argv = ['main.py', '--model', model_name, '--device', 'cuda', '--cuda', '0']
    from config.configurator import configs            # parses the unmodified YAML at import (configurator.py:57)
    yaml_model = dict(configs['model'])
    configs['train']['epoch'] = 2                      # a 500-epoch run is not a unit test; every other key is the YAML's
    import numpy as np
    import torch
    from data_utils.build_data_handler import build_data_handler
    from models.bulid_model import build_model
    from trainer.build_trainer import build_trainer
    from trainer.logger import Logger
    from trainer.trainer import init_seed

    init_seed()
    data_handler = build_data_handler()
    data_handler.load_data()
    model = build_model(data_handler).to(configs['device'])
    logger = Logger(log_configs=False)
    trainer = build_trainer(data_handler, logger)
    out = {'model_class': type(model).__module__ + '.' + type(model).__name__, 'trainer_class': type(trainer).__module__ + '.' + type(trainer).__name__,
           'yaml_model': {k: v for k, v in yaml_model.items() if isinstance(v, (int, float, str))},
           'state_dict_keys': sorted(model.state_dict().keys()), 'device': str(next(model.parameters()).device)}

    # ---- step 0 against the oracle on the reference's own first batch (deterministic terms) ----
    sys.path.insert(0, os.path.join(repo, 'tests'))
    from oracle import cf_oracle as O
    loader = data_handler.train_dataloader
    loader.dataset.sample_negs()
    torch.manual_seed(7)
    batch = next(iter(loader))
    batch_dev = [x.long().to(configs['device']) for x in batch]
    model.train()
    loss, parts = model.cal_loss(batch_dev)
    out['step0'] = {'loss': float(loss.detach()), **{k: float(torch.as_tensor(v).detach()) for k, v in parts.items()}}
    trn = data_handler.trn_mat
    adj = O.normalized_adjacency(trn.row.astype(np.int64), trn.col.astype(np.int64), trn.shape[0], trn.shape[1])
    ue, ie = model.user_embeds.detach().cpu().clone(), model.item_embeds.detach().cpu().clone()
    ancs, poss, negs = [x.long() for x in batch[:3]]
    mc = configs['model']
    n_layer = max(mc['layer_num'], 2 * mc['high_order']) if model_name == 'ncl' else mc['layer_num']
    x, layers = torch.cat([ue, ie]), []
    a_t = adj.torch_coo()
    layers.append(x)
    for _ in range(n_layer):
        layers.append(O.propagate(a_t, layers[-1]))
    e_clean = sum(layers[:mc['layer_num'] + 1])
    eu, ei = e_clean[:trn.shape[0]], e_clean[trn.shape[0]:]
    oracle = {'reg_loss': float(mc['reg_weight'] * O.reg_sumsq(list(p.detach().cpu() for p in model.parameters())))}
    if model_name in ('simgcl', 'sgl', 'ncl') or (model_name == 'lightgcn' and mc['keep_rate'] == 1.0):
        # BPR runs on the un-augmented propagation for these (simgcl.py:43,48; sgl.py:50-56; ncl.py:76-82)
        oracle['bpr_loss'] = float(O.bpr_loss_sum(eu[ancs], ei[poss], ei[negs]) / ancs.shape[0])
    out['oracle_step0'] = oracle
    loss.backward()
    out['grad_finite'] = bool(all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None))
    model.zero_grad()

    # ---- the reference's Trainer: train_epoch x2, evaluate (Metric.eval + torch.topk), test ----
    trainer.create_optimizer(model)
    before = model.user_embeds.detach().clone()
    for epoch in range(configs['train']['epoch']):
        trainer.train_epoch(model, epoch)
    out['params_moved'] = float((model.user_embeds.detach() - before).abs().max())
    res = trainer.evaluate(model, 0)
    out['reference_metric_eval'] = {k: [float(x) for x in v] for k, v in res.items()}
    res_t = trainer.test(model)
    out['reference_metric_test'] = {k: [float(x) for x in v] for k, v in res_t.items()}
    # the same model through this repository's evaluator (device mask + native top-k): identical numbers expected
    from sslrec_b200.trainer import Trainer as OurTrainer
    ours = OurTrainer(data_handler).evaluate(model, loader=data_handler.valid_dataloader)
    out['native_eval'] = {k: [float(x) for x in v] for k, v in ours.items()}
    # early-stop round trip of the reference (trainer.py:118,130-131): deepcopy(state_dict()) into a fresh build_model
    from copy import deepcopy
    sd = deepcopy(model.state_dict())
    fresh = build_model(data_handler).to(configs['device'])
    fresh.load_state_dict(sd)
    out['state_dict_roundtrip'] = bool(all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), sd.values())))
    from sslrec_b200 import _lib
    out['native_launches'] = _lib.launch_count()
    print('DROPIN_JSON ' + json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
