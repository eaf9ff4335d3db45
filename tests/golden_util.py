"""Load tests/golden/*.npz (written by oracle/make_golden.py from the reference's own source)."""
import functools
import os

import numpy as np
import torch

from oracle import egt_oracle as O


@functools.lru_cache(maxsize=None)
def _blob(golden_dir, kind):
    return np.load(os.path.join(golden_dir, f'egt_{kind}_reference.npz'), allow_pickle=False)


def n_cases(golden_dir, kind):
    return int(_blob(golden_dir, kind)['n_cases'])


def _case_items(golden_dir, kind, idx):
    z = _blob(golden_dir, kind)
    pre = f'case{idx:02d}/'
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def layer_case(golden_dir, idx, dtype=torch.float64):
    it = _case_items(golden_dir, 'layer', idx)
    flags = {}
    for k, v in it.items():
        if not k.startswith('flag_'):
            continue
        name = k[5:]
        if name == 'clip':
            flags['clip_logits_value'] = None if np.isnan(v).all() else tuple(float(x) for x in v)
        elif name == 'scaler_type':
            flags['scaler_type'] = str(v)
        elif v.dtype == np.bool_:
            flags[name] = bool(v)
        elif np.issubdtype(v.dtype, np.integer):
            flags[name] = int(v)
        else:
            flags[name] = float(v)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    c = dict(flags=flags, training=bool(int(it['training'])), mask=torch.from_numpy(it['mask']))
    for k in ('QKV', 'E', 'G', 'M', 'V_att', 'H_hat', 'A_tild'):
        c[k] = t(it[k])
    for k in ('uniform_noise', 'dropout_noise'):
        if k in it:
            c[k] = t(it[k])
    ins = [c['QKV']]
    if flags['edge_input']:
        ins.append(c['E'])
    if flags['gate_input']:
        ins.append(c['G'])
    if flags['attn_mask']:
        ins.append(c['M'])
    c['inputs'] = ins
    return c


def block_case(golden_dir, idx, dtype=torch.float64):
    it = _case_items(golden_dir, 'block', idx)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    kw = {}
    for k, v in it.items():
        if not k.startswith('cfg_'):
            continue
        name = k[4:]
        if name == 'clip':
            kw['clip_logits_value'] = None if np.isnan(v).all() else tuple(float(x) for x in v)
        elif v.dtype.kind in 'US':
            s = str(v)
            kw[name] = None if s == 'None' else s
        elif v.dtype == np.bool_:
            kw[name] = bool(v)
        elif np.issubdtype(v.dtype, np.integer):
            kw[name] = int(v)
        else:
            kw[name] = float(v)
    c = dict(cfg=O.BlockConfig(**kw), training=bool(int(it['training'])), mask=torch.from_numpy(it['mask']),
             params={k[6:]: t(v) for k, v in it.items() if k.startswith('param:')})
    for k in ('h', 'e', 'h_out', 'e_out', 'H_hat', 'V_att', 'h_ffn', 'e_ffn', 'edge_mask', 'uniform_noise'):
        if k in it:
            c[k] = t(it[k])
    return c
