"""The global ``configs`` dict of the reference (config/configurator.py:57), shared rather than
copied: inside an SSLRec checkout (``config.configurator`` already imported by main.py) this IS the
reference's dict, so ``config/modelconf/*.yml`` files drop in unchanged; standalone, ``load_config``
fills it from a YAML file with the same defaults the reference injects (configurator.py:26-51)."""
from __future__ import annotations

import sys

_ref = sys.modules.get('config.configurator')
configs: dict = _ref.configs if _ref is not None and hasattr(_ref, 'configs') else {}


def load_config(yaml_path: str = None, dataset: str = None, device: str = 'cuda', overrides: dict = None, base: dict = None) -> dict:
    """Fill ``configs`` in place (modules that did ``from sslrec_b200.config import configs`` see it)."""
    if base is not None:
        new = base
    else:
        import yaml
        with open(yaml_path, encoding='utf-8') as f:
            new = yaml.safe_load(f.read())
    new['model']['name'] = new['model']['name'].lower()
    new.setdefault('tune', {'enable': False})
    new['device'] = device
    if dataset is not None:
        new['data']['name'] = dataset
    new['train'].setdefault('log_loss', True)
    if 'patience' in new['train']:
        if new['train']['patience'] <= 0:
            raise Exception("'patience' should be greater than 0.")
        new['train']['early_stop'] = True
    else:
        new['train']['early_stop'] = False
    for section, kv in (overrides or {}).items():
        new.setdefault(section, {}).update(kv)
    configs.clear()
    configs.update(new)
    return configs


def default_config(model: str, **model_hp) -> dict:
    """The five in-scope YAMLs' common skeleton (lightgcn.yml:1-33 etc.) for programmatic use."""
    hp = dict(name=model, keep_rate=1.0, layer_num=2, reg_weight=1.0e-8, embedding_size=32)
    hp.update(model_hp)
    return {
        'optimizer': {'name': 'adam', 'lr': 1.0e-3, 'weight_decay': 0},
        'train': {'epoch': 1, 'batch_size': 4096, 'save_model': False, 'loss': 'pairwise', 'log_loss': False,
                  'test_step': 3, 'reproducible': True, 'seed': 2023},
        'test': {'metrics': ['recall', 'ndcg'], 'k': [10, 20, 40], 'batch_size': 1024},
        'data': {'type': 'general_cf', 'name': 'synthetic'},
        'model': hp,
    }
