"""CPU-side checks: the C-ABI library loads and exports every symbol include/sgqn_b200.h declares with the
argument lists the ctypes binding uses; parameter layout round-trips; lazy logging arithmetic."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
import sgqn_carla_b200 as S
from sgqn_carla_b200 import _lib
from sgqn_carla_b200.layout import ParamLayout, reference_key_map, FEAT


def _header_protos():
    src = open(os.path.join(ROOT, "include", "sgqn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(sgqn_\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        protos[m.group(1)] = args
    return protos


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    protos = _header_protos()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.sgqn_abi_version() == _lib.ABI_VERSION


def test_ctypes_signatures_match_header():
    import ctypes as C
    protos = _header_protos()
    assert set(_lib.SIGNATURES) == set(protos) - {"sgqn_abi_version"}
    for name, args in protos.items():
        if name == "sgqn_abi_version":
            continue
        sig = _lib.SIGNATURES[name]
        assert len(sig) == len(args), name
        for ct, a in zip(sig, args):
            if "*" in a:
                want = C.c_void_p
            elif a.startswith("unsigned long long"):
                want = C.c_ulonglong
            elif a.startswith("long long"):
                want = C.c_longlong
            elif a.startswith("float"):
                want = C.c_float
            elif a.startswith("double"):
                want = C.c_double
            else:
                assert a.startswith("int"), (name, a)
                want = C.c_int
            assert ct is want, (name, a)


def test_p2p_arena_header_layout():
    """The header the peer-memory collectives expect in front of the gradient arena (csrc/p2p.cu): a host-side query, no GPU."""
    import ctypes as C
    lay = (C.c_longlong * 3)()
    _lib.K.p2p_layout(lay)
    flags, ctl, small = list(lay)
    assert flags == 8 * 160 * 8 * 4 and ctl == 8 * 16 and small == 8 * 2 * 8 * 128      # slots x CTAs x ranks, slots x 4 words, slots x parity x ranks x 128 B
    assert all(v % 16 == 0 for v in (flags, ctl, small))
    with pytest.raises(_lib.KernelError):
        _lib.K.p2p_layout(None)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsgqn_b200.so")
    with pytest.raises(_lib.KernelError):
        _lib.load()


def test_engine_refuses_cpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        S.make_agent((9, 84, 84), (2,), S.default_args())


def test_layout_roundtrip_and_ranges():
    lay = ParamLayout(2)
    g = torch.Generator().manual_seed(0)
    params = {n: torch.randn(*shape, generator=g) for n, (_, _, shape) in lay.entries.items()}
    flat = torch.zeros(lay.total)
    lay.pack(params, flat)
    back = lay.unpack(flat)
    for n in params:
        assert torch.equal(back[n], params[n]), n
    for n, (o, st, _) in lay.entries.items():
        assert o % 4 == 0, n                                    # 16-byte aligned segments
    c0, c1 = lay.ranges["critic"]
    crit = sum(int(np.prod(s)) for n, (_, _, s) in lay.entries.items() if n.startswith(("Q1.", "Q2.", "cnn.", "critic_proj.")))
    assert crit == 3818798
    assert c1 - c0 >= crit and lay.ranges["Q2"][0] - lay.ranges["Q1"][0] == lay.q_stride
    # NHWC permutation of the projection columns: feature (c,y,x) -> (y,x,c)
    w = params["critic_proj.0.weight"]
    o = lay.off("critic_proj.0.weight")
    stored = flat[o:o + 100 * FEAT].reshape(100, 441, 32)
    assert torch.equal(stored[7, 5, 3], w[7, 3 * 441 + 5])
    keys = reference_key_map()
    extra = {n for n in keys if n.startswith(("curl.", "pad_", "soda_"))}
    assert set(keys) - extra == set(lay.entries)                 # CURLHead.W / the PAD head only exist in their algorithm's layout ...
    curl = ParamLayout(2, algorithm="curl")
    assert set(curl.entries) - set(lay.entries) == {"curl.W"}
    pad = ParamLayout(2, algorithm="pad")
    assert set(pad.entries) - set(lay.entries) == {n for n in extra if n.startswith("pad_")}
    assert pad.ranges["aux"][0] <= pad.off("pad_mlp.4.bias") < pad.ranges["aux"][1]
    soda = ParamLayout(2, algorithm="soda")
    assert set(soda.entries) - set(lay.entries) == {n for n in extra if n.startswith("soda_")}
    assert soda.ranges["soda"][1] == soda.ranges["aux"][1] and (soda.ranges["soda"][1] - soda.ranges["soda"][0]) % 4 == 0
    x0, x1 = curl.ranges["aux"]                                  # ... inside the range its optimiser owns (curl.py:16-20)
    assert x0 <= curl.off("curl.W") < x1 and x0 == curl.off("cnn.0.weight")


def test_lazy_scalar_arithmetic_without_device():
    from sgqn_carla_b200.lazylog import LazyScalar

    class FakeRing:
        def read(self, slot, serial, col):
            return [1.5, 2.5][slot]

    a = LazyScalar([(FakeRing(), 0, 0, 0, 1.0)])
    b = LazyScalar([(FakeRing(), 1, 1, 0, 1.0)])
    s = 0
    s += a
    s += b
    assert isinstance(s, LazyScalar) and abs(s / 2 - 2.0) < 1e-12 and abs(float(a) - 1.5) < 1e-12
    assert "%.04f" % s == "4.0000"


def test_default_args_match_reference_defaults():
    a = S.default_args()
    assert (a.batch_size, a.hidden_dim, a.projection_dim, a.num_shared_layers) == (128, 1024, 100, 11)
    assert (a.critic_tau, a.encoder_tau, a.aux_lr, a.alpha_blending, a.sgqn_quantile) == (0.01, 0.05, 3e-4, 0.2, 0.5)
    assert S.default_args(algorithm="rad").image_size == 100


@pytest.mark.ref
def test_default_args_against_reference_argparse():
    from oracle import ref_shim as R
    ref = vars(R.parse_args([]))
    mine = vars(S.default_args())
    for k, v in mine.items():
        if k in ref:
            assert ref[k] == v, k


def test_overlay_dataset_loaders(tmp_path):
    """the reference's on-disk overlay formats (datasets/carla/*.npy, utils.py:325-327; Places365 ImageFolder,
    augmentations.py:17-62) load into pools of the shapes / ranges the overlay kernels expect."""
    import importlib
    D = importlib.import_module("sgqn_carla_b200.datasets")
    rs = np.random.RandomState(0)
    carla = tmp_path / "carla"; carla.mkdir()
    frames = rs.randint(0, 256, size=(5, 3, 84, 84), dtype=np.uint8)
    for i, f in enumerate(frames):
        np.save(carla / f"frame_{i:04d}.npy", f)
    (carla / "notes.txt").write_text("ignored")
    got = D.load_carla_frames(str(carla))
    assert got.dtype == np.uint8 and got.shape == (5, 3, 84, 84) and np.array_equal(got, frames)
    assert D.load_carla_frames(str(carla), limit=2).shape[0] == 2
    np.save(carla / "zz_bad.npy", np.zeros((84, 84, 3), dtype=np.uint8))
    with pytest.raises(ValueError):
        D.load_carla_frames(str(carla))
    with pytest.raises(FileNotFoundError):
        D.load_carla_frames(str(tmp_path))
    # Places365 layout: <dir>/places365_standard/train/<class>/*.png
    from PIL import Image
    root = tmp_path / "data"
    for cls in ("a", "b"):
        d = root / "places365_standard" / "train" / cls; d.mkdir(parents=True)
        for i in range(3):
            Image.fromarray(rs.randint(0, 256, size=(100, 120, 3), dtype=np.uint8)).save(d / f"{i}.png")
    pool = D.load_places_pool([str(tmp_path / "missing"), str(root)], n=8, seed=1)
    assert pool.shape == (8, 3, 84, 84) and pool.dtype == torch.float32
    assert float(pool.min()) >= 0.0 and float(pool.max()) <= 1.0 and float(pool.std()) > 0.05
    assert torch.equal(pool, D.load_places_pool(str(root), n=8, seed=1))                # reproducible
    assert not torch.equal(pool, D.load_places_pool(str(root), n=8, seed=2))
    with pytest.raises(FileNotFoundError):
        D.load_places_pool([str(tmp_path / "missing")], n=2)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: rank 0 prints ONE JSON line with the contract's keys (the oracle port of the reference update on
    the host cores); other ranks exit 0 without output."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "updates/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("SGSAC updates/sec")
