"""CPU: the C-ABI library loads and exports every symbol include/mt_b200.h declares, the Python structs match the C
layout, the drop-in modules have the reference's state_dict keys, and nothing falls back to the CPU."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import _lib
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODS = ['acoustic', 'image', 'linguistic']
DIMS = {'acoustic': 88, 'image': 256, 'linguistic': 300}


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'mt_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return set(re.findall(r'\b(mt_[a-z0-9_]+)\s*\(', src))


def test_header_and_binding_agree():
    assert header_symbols() == set(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()                       # raises if the .so is missing or a symbol is absent
    for name in header_symbols():
        assert hasattr(L, name), name
    assert L.mt_version() >= 100
    assert L.mt_error_string(0) == b'ok' and b'workspace' in L.mt_error_string(4)


def test_struct_layouts_and_size_queries():
    L = _lib.lib()
    # parameter counts are pure host arithmetic: compare with the reference inventory
    inv = util.key_inventory()['MFT.MultiTransformer']
    enc = sum(int(torch.tensor(s).prod()) for k, s in inv.items() if k.startswith('transformer_acoustic.'))
    assert L.mt_encoder_param_count(256, 128, 6) == enc
    cfg = _lib.MtMfnCfg()
    cfg.B, cfg.T, cfg.n_mods = 4, 8, 3
    for i, (d, h) in enumerate([(256, 48), (256, 88), (256, 88)]):
        cfg.in_dim[i] = d; cfg.hid[i] = h
    cfg.mem_dim, cfg.h_att1, cfg.h_att2, cfg.h_gamma, cfg.h_out = 128, 128, 256, 64, 64
    mfn = sum(int(torch.tensor(s).prod()) for k, s in inv.items() if k.startswith('mfn.'))
    assert L.mt_mfn_param_count(ctypes.byref(cfg)) == mfn
    assert L.mt_mfn_ws_bytes(ctypes.byref(cfg)) > 0
    ecfg = _lib.MtEncoderCfg(2, 16, 256, 8, 128, 6, 0, 1, 0.1, 0, 0, 1, 0, None)
    tr = L.mt_encoder_ws_bytes(ctypes.byref(ecfg))
    ecfg.training = 0
    assert 0 < L.mt_encoder_ws_bytes(ctypes.byref(ecfg)) < tr
    ecfg.d = 250                          # not a multiple of 128 -> unsupported, reported as 0 bytes
    assert L.mt_encoder_ws_bytes(ctypes.byref(ecfg)) == 0
    hcfg = _lib.MtLstmHeadCfg(2, 8, 256, 128, 0, 1)
    sft = util.key_inventory()['SFT.NLPTransformer']
    dec = sum(int(torch.tensor(s).prod()) for k, s in sft.items() if k.startswith(('decoder.', 'dec_', 'out.')))
    assert L.mt_lstm_head_param_count(ctypes.byref(hcfg)) == dec


def test_null_and_bad_arguments_are_errors_not_crashes():
    L = _lib.lib()
    assert L.mt_layernorm_fwd(0, 4, 256, None, None, None, 1e-6, None, 1, None) == 1
    assert L.mt_attention_fwd(0, 1, 4, 250, 8, None, None, None, None, 0.0, 0, 0, None) != 0
    with pytest.raises(RuntimeError):
        _lib.check(L.mt_adam_step(None, None, None, None, 4, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None))


@pytest.mark.parametrize('inv_name,ctor', [
    ('MFT.MultiTransformer', lambda: mtb.MultiTransformer(MODS, DIMS)),
    ('B3.MultiTransformer', lambda: mtb.B3MultiTransformer(MODS, {'acoustic': 256, 'image': 256, 'linguistic': 300})),
    ('SFT.NLPTransformer', lambda: mtb.NLPTransformer(512)),
    ('MFT.UniFullTransformer', lambda: mtb.UniFullTransformer(556)),
    ('MFT.UniTransformer', lambda: mtb.UniTransformer(300)),
])
def test_state_dict_keys_match_reference(inv_name, ctor):
    inv = util.key_inventory()[inv_name]
    sd = ctor().state_dict()
    assert list(sd.keys()) == list(inv.keys())
    for k, s in inv.items():
        assert list(sd[k].shape) == s, k


def test_reference_checkpoint_roundtrip():
    m = mtb.MultiTransformer(MODS, DIMS, N=2)
    sd = util.filled_sd(util.mods_shapes('MFT.MultiTransformer', 2), 7)
    m.load_state_dict(sd)
    out = m.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    # the orphan attn{mod}/ff{mod} templates exist (checkpoint compat) and 987,264 parameters are dead
    full = mtb.MultiTransformer(MODS, DIMS)
    dead = sum(p.numel() for n, p in full.named_parameters() if n.startswith(('attn', 'ff')))
    assert dead == 987264
    assert sum(p.numel() for p in full.parameters()) == 7775041


def test_no_cpu_fallback():
    m = mtb.MultiTransformer(MODS, DIMS, N=1)
    x = {k: torch.zeros(2, 4, d) for k, d in DIMS.items()}
    with pytest.raises(RuntimeError, match='CUDA'):
        m(x, torch.ones(2, 4, 1), [4, 4])
    with pytest.raises(RuntimeError, match='CUDA'):
        mtb.LayerNorm(256)(torch.zeros(2, 256))


def test_arena_binding_preserves_parameters_cpu_logic():
    from multimodal_transformer_b200.functional import Arena
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    a = Arena(ps)
    assert a.total == 17 and a.offsets == [0, 12] and not a.bound()
    with pytest.raises(RuntimeError):
        a.bind()                          # CPU parameters: refuse rather than fall back


def test_no_product_import_of_oracle():
    out = subprocess.run(['grep', '-rlE', r'^\s*(from|import)\s+oracle', os.path.join(ROOT, 'multimodal-transformer_b200'),
                          os.path.join(ROOT, 'multimodal_transformer_b200')], capture_output=True, text=True).stdout
    assert out.strip() == ''


def test_encoder_rejects_unsupported_shapes_at_construction():
    """The reference's 'emotient' modality (d_model 16, h 8 -> d_k 2) is outside the sm_100a kernels: the constructor says so
     instead of failing at the first forward."""
    with pytest.raises(NotImplementedError, match='emotient'):
        mtb.MultiTransformer(['emotient', 'linguistic'], {'emotient': 20, 'linguistic': 300})
    with pytest.raises(NotImplementedError):
        mtb.Encoder(mtb.EncoderLayer(16, mtb.MultiHeadedAttention(8, 16), mtb.PositionwiseFeedForward(16, 128), 0.1), 2)
    mtb.MFN(['emotient', 'linguistic'], {'emotient': 16, 'linguistic': 256}, 1)      # the MFN alone takes it


def _load_bench():
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('mt_bench_module', os.path.join(root, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_bench_clock_summary_merges_nvml_and_nvidia_smi_samples():
    """bench.py's ClockSampler: NVML samples (10 ms) and nvidia-smi rows (200 ms) land in one record -- median SM clock over both,
    throttle reasons from either source (NVML bit masks of nvmlClocksEventReason*), counts per source; no sampler at all is said so."""
    b = _load_bench()
    cs = b.ClockSampler(0)
    cs.nvml_rows = [(1965.0, 0), (1965.0, 0x4), (1950.0, 0)]          # one sample under sw_power_cap
    cs.nvml_max = 1965.0
    out = cs._summary([1965.0], [1965.0], set())
    assert out['samples'] == 4 and out['samples_nvml'] == 3 and out['samples_nvidia_smi'] == 1
    assert out['sm_mhz'] == 1965.0 and out['sm_max_mhz'] == 1965.0
    assert out['reasons'] == ['sw_power_cap']
    cs.nvml_rows = [(1200.0, 0x40 | 0x8)]
    out = cs._summary([], [], {'sw_thermal_slowdown'})
    assert out['reasons'] == ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'] and out['sm_mhz'] == 1200.0
    cs2 = b.ClockSampler(0)
    cs2.start()                                                        # no GPU here: neither NVML nor nvidia-smi
    rec = cs2.stop()
    assert rec['sm_mhz'] is None or rec.get('samples', 0) >= 0


def test_bench_parses_its_contract_flags():
    """`python bench.py --gpus N --steps K --warmup W [--impl reference]` is the driver's contract; defaults are N = 1 and a K / W that
    finish within minutes; --sync-e2e keeps the drain-every-step e2e measurement available."""
    import sys
    b = _load_bench()
    argv, sys.argv = sys.argv, ['bench.py', '--gpus', '2', '--steps', '7', '--warmup', '4', '--impl', 'reference', '--sync-e2e']
    try:
        a = b.parse()
    finally:
        sys.argv = argv
    assert (a.gpus, a.steps, a.warmup, a.impl, a.sync_e2e) == (2, 7, 4, 'reference', True)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        d = b.parse()
    finally:
        sys.argv = argv
    assert d.gpus == 1 and d.impl == 'ours' and d.config == 'c2' and d.steps <= 50 and d.warmup >= 3 and not d.sync_e2e
