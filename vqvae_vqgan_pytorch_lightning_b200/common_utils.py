"""Config plumbing: YAML -> the dicts VQVAE takes (reference: vqvae/common_utils.py:30-35 get_model_conf and the
t_conf / batch-size / learning-rate derivation of vqvae/train.py:56-98)."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import yaml


def get_model_conf(filepath: str) -> dict:
    with open(filepath, 'r', encoding='utf-8') as f:
        return yaml.safe_load(f)


def derive_confs(conf: dict, world_size: int = 1, overrides: Optional[dict] = None) -> Tuple[int, dict, dict, Optional[dict], dict, int]:
    """-> (image_size, ae_conf, q_conf, l_conf, t_conf, batch_size_per_device).  lr = base_lr * sqrt(cumulative_bs/256)
    (train.py:62-63); per-device batch = cumulative_bs // world (train.py:59-60).  `overrides` may replace
    image_size / num_embeddings / cumulative_bs (BASELINE.json configs override the YAML values, SURVEY.md TL;DR)."""
    overrides = overrides or {}
    tr = dict(conf['training'])
    cumulative_bs = int(overrides.get('cumulative_bs', tr['cumulative_bs']))
    lr = float(tr['base_lr']) * math.sqrt(cumulative_bs / 256)
    q_conf = dict(conf['quantizer'])
    if 'num_embeddings' in overrides:
        q_conf['num_embeddings'] = int(overrides['num_embeddings'])
    t_conf = {'lr': lr, 'betas': tr['betas'], 'eps': tr['eps'], 'weight_decay': tr['weight_decay'],
              'warmup_epochs': tr.get('warmup_epochs'), 'decay_epochs': tr.get('decay_epochs')}
    image_size = int(overrides.get('image_size', conf['image_size']))
    return image_size, conf['autoencoder'], q_conf, conf.get('loss'), t_conf, cumulative_bs // max(1, world_size)
