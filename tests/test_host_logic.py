"""CPU: host-side logic of the drop-in layer and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, split_sd
from dpi_b200 import _lib, synthetic as syn


def _header_functions():
    text = open(os.path.join(ROOT, "include", "dpi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpi_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dpi_b200.h but not exported"
    # and the Python binding lists exactly the header's functions
    assert sorted(_lib.PROTOTYPES) == names
    assert _lib.load().dpi_version() >= 100


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _obs_from_fixture(g):
    """Rebuild the duck-typed obs from the numeric fixture (names are irrelevant to the maps)."""
    obs = syn.SyntheticObs(n_stations=int(g["n_stations"]), n_scans=int(g["n_scans"]), seed=int(g["seed"]),
                           min_vis=int(g["min_vis"]), flip_frac=0.4)
    assert np.array_equal(obs.data["time"], g["time"]) and np.allclose(obs.data["u"], g["u"])
    return obs


@pytest.mark.parametrize("name", ["closure_small", "closure_eht"])
def test_closure_index_maps_bit_exact(name):
    """Our vectorised Obs_params_torch vs the tables the reference's own loops produced
    (interferometry_helpers.py:94-228): int64 indices and fp64 signs must be identical."""
    from dpi_b200.interferometry_helpers import Obs_params_torch
    g = load_golden(name)
    obs = _obs_from_fixture(g)
    dft, ktraj, pf, cpi, cps, cai = Obs_params_torch(obs, syn.SimImage(int(g["npix"]), 160.0), ttype="direct")
    for q in range(3):
        assert cpi[q].dtype == torch.int64 and cps[q].dtype == torch.float64
        assert np.array_equal(cpi[q].numpy(), g["cp_ind"][q])
        assert np.array_equal(cps[q].numpy(), g["cp_sign"][q])
    for q in range(4):
        assert cai[q].dtype == torch.int64
        assert np.array_equal(cai[q].numpy(), g["ca_ind"][q])
    ref_dft = g["dft_mat"]
    assert dft.dtype == torch.float32
    assert np.array_equal(dft.numpy()[:, :ref_dft.shape[1]], ref_dft)
    assert np.array_equal(ktraj.numpy(), g["ktraj"]) and np.array_equal(pf.numpy(), g["pulsefac"])
    # both signs occur, and index 0 keeps its sign through the zero_symbol encoding
    assert (g["cp_sign"] == -1).any() and (g["cp_sign"] == 1).any()


def test_zero_index_sign_is_preserved():
    from dpi_b200.interferometry_helpers import _signed_map, _ind_sign
    t = np.array([0.0, 0.0, 0.0])
    t1, t2 = np.array(["B", "A", "A"]), np.array(["A", "C", "B"])   # row 0 stores (B, A): reversed
    m = _signed_map(t, t1, t2, np.array([0.0]), [(np.array(["A"]), np.array(["B"]))], 100000)
    # rows 0 (reversed) and 2 (forward) both match; the later row overwrites (reference loop order)
    assert m[0, 0] == 2
    m = _signed_map(t[:2], t1[:2], t2[:2], np.array([0.0]), [(np.array(["A"]), np.array(["B"]))], 100000)
    ind, sign = _ind_sign(m[:, 0], 100000)
    assert ind[0] == 0 and sign[0] == -1


def test_dft_nfft_branches_target_the_same_sum():
    """ttype='nfft' is served by the exact sum its NUFFT approximates; it must agree with ftmatrix."""
    from dpi_b200.interferometry_helpers import Obs_params_torch, dft_from_ktraj
    obs = syn.SyntheticObs(n_stations=5, n_scans=6, seed=1, min_vis=3)
    sim = syn.SimImage(16, 160.0)
    dft, ktraj, pf, *_ = Obs_params_torch(obs, sim, ttype="direct")
    dft2 = dft_from_ktraj(16, ktraj, pf)
    assert float((dft - dft2).abs().max()) < 2e-6
    img = syn.ground_truth_image(16, 2.0).reshape(-1)
    zero_bl = syn.ftmatrix(sim.psize, 16, 16, np.zeros((1, 2)))
    assert abs((zero_bl @ img)[0] - 2.0) < 1e-12  # zero baseline = total flux


@pytest.mark.parametrize("name", ["flow_even", "flow_odd", "flow_nobn"])
def test_state_dict_round_trip_with_reference_keys(name):
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    g = load_golden(name)
    sd = split_sd(g, "sd_after_rev.")
    m = RealNVP(int(g["D"]), int(g["NF"]), seqfrac=float(g["seqfrac"]), batch_norm=bool(g["batch_norm"]))
    fresh = m.state_dict()
    assert list(fresh.keys()) == list(sd.keys())          # same keys, same order as the reference module
    for k in sd:
        assert tuple(fresh[k].shape) == tuple(sd[k].shape) and fresh[k].dtype == sd[k].dtype, k
    res = m.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    back = m.state_dict()
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    # one flat parameter, laid out in the reference's parameters() order
    params = list(m.parameters())
    assert len(params) == 1 and params[0] is m.flat
    names = [k for k, _ in m.named_parameter_views()]
    assert names == [k for k in sd if not re.search(r"initialized|running|tracked", k)]
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "actnorm2.loc" not in k})


def test_fresh_module_matches_reference_init_distribution():
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    torch.manual_seed(0)
    m = RealNVP(64, 2)
    sd = m.state_dict()
    assert abs(float(sd["blocks.0.coupling.net.0.weight"].std()) - 0.05) < 0.01
    assert float(sd["blocks.0.coupling.net.6.fc.weight"].abs().max()) == 0.0
    assert float(sd["blocks.1.coupling2.net.2.weight"].min()) == 1.0
    assert float(sd["blocks.1.coupling2.net.5.running_var"].min()) == 1.0
    assert int(sd["blocks.0.actnorm.initialized"]) == 0
    for i in range(2):
        assert np.array_equal(m.orders[i], np.random.RandomState(seed=i).permutation(64))
        assert np.array_equal(m.orders[i][m.inverse_orders[i]], np.arange(64))


def test_product_path_has_no_cpu_fallback():
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    from dpi_b200 import interferometry_helpers as ih
    m = RealNVP(8, 1, seqfrac=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.reverse(torch.randn(4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ih.Loss_l1(torch.rand(2, 4, 4))
    # nothing under dpi_b200/ may import the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dpi_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


@pytest.mark.parametrize("M,N,K", [(1024, 512, 2048), (1024, 512, 512), (1024, 4096, 512), (1024, 512, 4096),
                                   (4096, 512, 1024), (512, 2048, 1024), (2048, 144, 9), (2048, 18, 144),
                                   (144, 144, 2048), (200, 32, 128), (97, 4, 5), (128, 64, 32)])
def test_tcgen05_gemm_planner_invariants(M, N, K):
    """flow_tc.cu's host planner (no CUDA call): tiles cover the problem, the K splits cover K exactly once, the grid
    is at most one wave of 148 CTAs whenever K is split, the stage ring fits the 227 KB of shared memory, the TMEM
    accumulator is a legal power-of-two column count."""
    lib = _lib.load()
    out = (ctypes.c_int32 * 10)()
    n_sm = 148
    _lib.check(lib.dpi_flow_tc_plan(M, N, K, n_sm, out), "dpi_flow_tc_plan")
    TN, KT, MH, m_tiles, n_tiles, k_tiles, splits, kt_per_split, stages, smem = list(out)
    assert (TN, KT, MH) in {(64, 32, 1), (128, 32, 1), (256, 32, 1), (256, 16, 1), (256, 16, 2)}
    assert m_tiles * 128 >= M and (m_tiles - MH) * 128 < M and m_tiles % MH == 0
    assert n_tiles * TN >= N and (n_tiles - 1) * TN < N
    assert k_tiles * KT >= K and (k_tiles - 1) * KT < K
    assert splits >= 1 and splits * kt_per_split >= k_tiles and (splits - 1) * kt_per_split < k_tiles
    ctas = (m_tiles // MH) * n_tiles * splits
    if splits > 1:
        assert ctas <= n_sm                       # never a second, nearly empty wave
        assert kt_per_split * KT >= 128           # at least 128 k per split
    assert 2 <= stages <= 6 and smem <= 227 * 1024
    cols = MH * TN
    assert cols in (64, 128, 256, 512)
    if N >= 256 and (M + 127) // 128 * ((N + 255) // 256) * min(16, max(1, K // 128)) >= 111:
        assert TN == 256                          # the widest tile that still fills three quarters of the GPU


def test_glow_module_surface_and_reference_state_dict():
    """the drop-in glow_model has the reference's constructor, calc_z_shapes and state_dict keys / shapes -
    a state dict written by the unmodified reference (tests/golden/glow_small.npz) loads with strict=True - and it
    refuses CPU tensors and gradient requests instead of falling back (no CUDA call is made here)."""
    import warnings
    from dpi_b200.generative_model.glow_model import Glow, calc_z_shapes
    from oracle import glow_oracle as GO
    g = load_golden("glow_small")
    sd = split_sd(g)
    flows = sorted({k[: k.index("coupling.")] for k in sd if "coupling.net.0.weight" in k})
    for tag, k in enumerate(GO.big_weight_keys([f + "coupling.net.3.weight" for f in flows])):
        sd[k] = GO.formula_weight((GO.FILTER, GO.FILTER, 1, 1), tag)
    m = Glow(1, int(g["n_flow"]), int(g["n_block"]))
    mine = m.state_dict()
    assert sorted(mine) == sorted(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape) and mine[k].dtype == sd[k].dtype, k
    m.load_state_dict(sd, strict=True)
    for npix, nf, nb in [(32, 16, 4), (8, 1, 2), (64, 4, 5)]:
        assert calc_z_shapes(1, npix, nf, nb) == GO.calc_z_shapes(1, npix, nf, nb)
    assert calc_z_shapes(1, 32, 16, 4) == [(2, 16, 16), (4, 8, 8), (8, 4, 4), (32, 2, 2)]
    assert sum(p.numel() for p in Glow(1, 16, 4).parameters()) == 23677792
    zs = [torch.randn(2, *s) for s in calc_z_shapes(1, 8, 1, 2)]
    with pytest.raises(RuntimeError, match="no CPU path"):                 # a training call on CPU tensors: no fallback
        m.reverse([z.clone().requires_grad_(True) for z in zs])
    m.eval()
    with pytest.raises(NotImplementedError, match="eval-mode"):            # gradients through running statistics: not built
        m.reverse(zs)
    m.train()
    with pytest.raises(NotImplementedError, match="training backward"):    # density direction: no backward, never silent
        m.forward(torch.randn(2, 1, 8, 8))
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU path"):
            m.reverse(zs)
        with pytest.raises(RuntimeError, match="no CPU path"):
            m.forward(torch.randn(2, 1, 8, 8))
    with pytest.raises(NotImplementedError):
        Glow(1, 1, 2, affine=False)


def test_native_handles_are_validated_and_never_shared_by_copies():
    """a raw handle duplicated by copy.deepcopy / pickle must not be double-freed, and an unknown pointer is
    rejected by the C ABI instead of being dereferenced (no CUDA call is made here)."""
    import copy
    import ctypes as C
    import pickle
    from dpi_b200 import _lib
    from dpi_b200.generative_model.realnvpfc_model import RealNVP
    lib = _lib.load()
    bogus = C.c_void_p(0xDEADBEE0)
    assert lib.dpi_flow_destroy(bogus) != 0 and b"destroyed" in lib.dpi_last_error()
    assert lib.dpi_eht_destroy(bogus) != 0
    assert lib.dpi_flow_param_count(bogus) == -1 and lib.dpi_flow_workspace_bytes(bogus, 8) == 0
    assert lib.dpi_flow_cl_active(bogus, 8) == 0 and lib.dpi_eht_workspace_bytes(bogus, 8) == 0
    assert lib.dpi_flow_run(bogus, 0, 1, 8, None, None, None, None, None, None, 0, None) != 0
    m = RealNVP(8, 2, seqfrac=1)
    m._handles[0] = 0xDEADBEE0                      # stands for a live native handle
    for other in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert other._handles == {} and other._last_ws is None
        assert torch.equal(other.flat.detach(), m.flat.detach())
        assert other.blocks[1].coupling2 is not None and other.blocks[0]._m is other
    m._handles.clear()
