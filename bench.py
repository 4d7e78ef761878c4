#!/usr/bin/env python
"""SGSAC updates/sec on B200 (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N=1 workload = BASELINE.json configs[1]: SGSAC full update loop (critic + 2 attributions + mask-consistency,
actor/alpha, target EMA, overlay aux update; all frequencies 2 so steps alternate odd/even), batch 128,
9x84x84 uint8 frame stacks, synthetic replay (SURVEY.md 8d cfg 2).  N>1: the batch is sharded, 128 samples per
rank (global batch 128*N = BASELINE config 4 at N=8), gradients all-reduced with NCCL; weak scaling.

`value`   : updates/s with the replay ring resident in HBM (CUDA-event timed, max over ranks), times
            global_batch/128 so it is the whole-job aggregate in batch-128 updates.
`e2e`     : the same through the public API with HOST buffers: the replay frame ring lives in pinned host
            memory and every step's sampled frames cross to the device inside the timed region, and the
            step's loss vector is read back to the host every step.
`roofline`: the dominant kernel family of the step, timed live with CUDA events around its launches.
`cpu_baseline`: the oracle port of the reference update (oracle/sgsac_oracle.py) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL's own log (communicator sizes, NVLS / ring choice) goes to stderr with everything else a library prints: stdout is
# redirected to stderr while the benchmark runs (_StdoutToStderr) and carries exactly the one JSON line
os.environ["NCCL_DEBUG"] = os.environ.get("SGQN_NCCL_DEBUG", "INFO" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "WARN")

import numpy as np  # noqa: E402
import torch  # noqa: E402

PER_GPU_BATCH = 128
CAPACITY = 20000            # transitions; 20003 frames x 21 KB = 423 MB > 126 MB L2
POOL_N = 2048               # overlay frames (43 MB)
# algorithmic FLOPs per sample (minimal / de-duplicated schedule), FLOP = 2*MAC
GFLOP_ODD, GFLOP_EVEN = 1.671, 3.712
WORKLOAD = ("SGSAC full update loop (critic + attribution mask consistency + actor/alpha + target EMA + overlay aux), "
            "batch 128 per GPU, 9x84x84 uint8 stacks, A=2, sgqn_quantile=0.95, reference init, steps alternate odd/even")


def env_int(k, d):
    return int(os.environ.get(k, d))


class NullLog:
    def __init__(self):
        self.last = {}

    def log(self, k, v, step, n=1):
        self.last[k] = v


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def synthetic(capacity, A=2, seed=0):
    rs = np.random.RandomState(seed)
    frames = rs.randint(0, 256, size=(capacity + 3, 3, 84, 84), dtype=np.uint8)
    actions = rs.uniform(-1, 1, size=(capacity, A)).astype(np.float32)
    rewards = rs.randn(capacity, 1).astype(np.float32)
    not_dones = np.ones((capacity, 1), dtype=np.float32)
    pool = rs.randint(0, 256, size=(POOL_N, 3, 84, 84), dtype=np.uint8)
    return frames, actions, rewards, not_dones, pool


def synthetic_sized(capacity, A, size, seed=0):
    rs = np.random.RandomState(seed)
    frames = rs.randint(0, 256, size=(capacity + 3, 3, size, size), dtype=np.uint8)
    actions = rs.uniform(-1, 1, size=(capacity, A)).astype(np.float32)
    rewards = rs.randn(capacity, 1).astype(np.float32)
    not_dones = np.ones((capacity, 1), dtype=np.float32)
    pool = rs.randint(0, 256, size=(512, 3, 84, 84), dtype=np.uint8)
    return frames, actions, rewards, not_dones, pool


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def time_oracle(steps, warmup, budget_s, B=PER_GPU_BATCH):
    """Times the oracle port of SGSAC.update on the host cores (config 1 of BASELINE.json)."""
    from oracle import sgsac_oracle as O
    from oracle.pin_rnd import make_rnd
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    args = O.Args(algorithm="sgsac", sgqn_quantile=0.95, batch_size=B)
    rs = np.random.RandomState(0)
    rep = O.synthetic_replay(1000, 2, seed=0)
    pool = torch.as_tensor(rs.randint(0, 256, size=(256, 3, 84, 84), dtype=np.uint8))
    orc = O.make_oracle((9, 84, 84), (2,), args, seed=0)
    orc.pool = pool
    L = NullLog()
    step = 1
    for _ in range(warmup):
        orc.update_from_batch(rep.sample(rs.randint(0, 1000, size=B)), make_rnd(rs, B, 2, 256), L, step); step += 1
    times = []
    t_start = time.time()
    for i in range(steps):
        t0 = time.time()
        orc.update_from_batch(rep.sample(rs.randint(0, 1000, size=B)), make_rnd(rs, B, 2, 256), L, step); step += 1
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s and len(times) >= 2 and len(times) % 2 == 0:
            break
    odd = [t for i, t in enumerate(times) if (warmup + 1 + i) % 2 == 1]
    even = [t for i, t in enumerate(times) if (warmup + 1 + i) % 2 == 0]
    mean = 0.5 * (np.mean(odd) + np.mean(even)) if odd and even else float(np.mean(times))
    return 1.0 / mean, cores, f"{len(times)} consecutive oracle-port updates (B={B}, config 1, {len(even)} even + {len(odd)} odd) after {warmup} warm-up"


def run_reference(a):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    v, cores, sample = time_oracle(a.steps, min(a.warmup, 2), 150.0)
    line = {"impl": "reference", "metric": "SGSAC updates/sec (batch 128, 9x84x84)", "value": v, "unit": "updates/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH, "parallelism": "host-cpu",
                       "note": "oracle port of the reference's SGSAC.update (bit-exact with the reference in the build container) on the host cores"},
            "cpu_baseline": {"value": v, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- B200 arm
def profile_kernels(agent, rb, nsteps=4):
    """Per-ABI-call CUDA-event timing over `nsteps` updates -> {family: (total_ms, calls)}."""
    from sgqn_carla_b200 import _lib
    api = _lib.K
    recs = []
    names = list(_lib.SIGNATURES.keys())
    saved = {}
    for full in names:
        n = full[len("sgqn_"):]
        fn = getattr(api, n)
        saved[n] = fn

        def wrap(*a0, _fn=fn, _n=n):
            args = a0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if _n == "conv_chain":                      # the layer table only lives during the call: summarise it now
                import ctypes
                arr = (_lib.ConvLayer * args[1]).from_address(args[0])
                args = ("chain", [(L.B, L.Hr, L.Wp, L.Hv, L.shift, (L.flags >> 2) & 3) for L in arr])
            e0.record()
            _fn(*a0)
            e1.record()
            recs.append((_n, args, e0, e1))
        setattr(api, n, wrap)
    graphs = agent.use_cuda_graphs
    agent.use_cuda_graphs = False                     # per-call timing needs the eager path
    overlap = agent.engine.overlap
    agent.engine.overlap = False                      # ... and every kernel on the stream the events are recorded on
    try:
        L = NullLog()
        for s in range(1, nsteps + 1):
            # park the stream behind a ~40 ms spin so that the host (python + ctypes + tensor-map encoding, ~10 us per call)
            # has queued the whole update before the GPU starts on it: the event pairs then bracket device time only
            torch.cuda.synchronize()
            torch.cuda._sleep(int(0.04 * 1.9e9))
            agent.update(rb, L, s)
        torch.cuda.synchronize()
    finally:
        agent.use_cuda_graphs = graphs
        agent.engine.overlap = overlap
        for n, fn in saved.items():
            setattr(api, n, fn)
    # the dominant kernel as the product runs it: the conv_tc launches of ONE update, in program order, re-issued back to back
    # on one stream inside a CUDA graph (programmatic dependent launch overlaps each launch's prologue with its predecessor's
    # last tiles, layer l+1 reads layer l's output from L2) and timed with CUDA events around replays of that graph
    ingraph = None
    conv_calls = [(n, args) for n, args, _, _ in recs if n in ("conv_tc", "conv_chain")]
    conv_calls = conv_calls[-(len(conv_calls) // nsteps) * 2:]          # the last two updates = one odd + one even step
    if conv_calls and conv_calls[0][0] == "conv_tc":
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            for n, args in conv_calls:
                saved[n](*args[:-1], st)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        by = sum(128.0 * a[6] * ((a[9] + 2) * (a[9] + 2) + a[9] * a[9] * (2 if (a[-2] >> 2) & 3 else 1)) for _, a in conv_calls)
        fl = sum(2.0 * a[6] * a[9] * a[9] * 9 * 32 * 32 for _, a in conv_calls)
        ingraph = {"launches": len(conv_calls), "us_per_launch": e0.elapsed_time(e1) * 1e3 / (reps * len(conv_calls)),
                   "bytes_per_launch": by / len(conv_calls), "flops_per_launch": fl / len(conv_calls)}
    dump = os.environ.get("SGQN_PROFILE_CALLS")
    if dump:
        rows = [[n, [a for a in args if isinstance(a, int) and abs(a) < (1 << 31)][:14], round(e0.elapsed_time(e1), 4)] for n, args, e0, e1 in recs]
        json.dump(rows, open(dump, "w"))
    fam = {}
    for n, args, e0, e1 in recs:
        key = n
        if n in ("conv_fwd", "conv_dgrad", "conv_wgrad"):
            cin, cout = (args[7], args[8])
            key = f"{n}[{cin}->{cout}]"
        t = e0.elapsed_time(e1)
        fl, by = 0.0, 0.0
        if n == "conv_fwd":
            B, Hs, Ws, Cin, Cout, pad, up = args[4], args[5], args[6], args[7], args[8], args[9], args[10]
            Ho = Hs * up + 2 * pad - 2
            fl = 2.0 * B * Ho * Ho * 9 * Cin * (9 if Cout == 16 else Cout)
        elif n == "conv_dgrad":
            B, Hl, Wl, Cin, Cout, pad = args[4], args[5], args[6], args[7], args[8], args[9]
            Ho = Hl + 2 * pad - 2
            fl = 2.0 * B * Ho * Ho * 9 * Cin * (9 if Cout == 16 else Cout)
        elif n == "conv_tc":
            B, Hv, masked = args[6], args[9], (args[-2] >> 2) & 3
            fl = 2.0 * B * Hv * Hv * 9 * 32 * 32
            # algorithmic bytes: every input pixel read once, every output pixel written once, 128 B each;
            # the data gradient also reads the 128-byte ReLU mask (the layer's input activation) of every output pixel
            by = 128.0 * B * ((Hv + 2) * (Hv + 2) + Hv * Hv * (2 if masked else 1))
            key = "conv_tc[32->32 " + ("dgrad" if args[11] else "fwd") + "]"
        elif n == "conv_chain":
            # one persistent launch = the ten 32->32 layers of an encoder pass (or their data gradients): the same algorithmic
            # bytes / FLOPs as the ten per-layer launches it replaces
            masked = args[1][0][5] != 0
            for (B, Hr, Wp, Hv, shift, mm) in args[1]:
                fl += 2.0 * B * Hv * Hv * 9 * 32 * 32
                by += 128.0 * B * ((Hv + 2) * (Hv + 2) + Hv * Hv * (2 if mm else 1))
            key = "conv_chain[32->32 x10 " + ("dgrad" if args[1][0][4] else "fwd") + "]"
        elif n == "conv1_fused_tc":
            fl = 2.0 * args[5] * 1681 * 81 * 32
        elif n == "conv_tcg_taps":
            B, Cin, Cout, Hv = args[5], args[8], args[9], args[10]
            fl = 2.0 * B * Hv * Hv * 81 * 32
            key = "conv1_tcg[dgrad]"
        elif n == "conv_tcg":
            B, Cin, Cout, Hv = args[5], args[8], args[9], args[10]
            fl = 2.0 * B * Hv * Hv * 9 * Cin * Cout
            key = f"conv_tcg[{Cin}->{Cout}]"
        elif n == "conv_wgrad_tcg":
            B, Hr, Cin, Cout = args[3], args[4], args[6], args[7]
            fl = 2.0 * B * (Hr - 2) * (Hr - 2) * 9 * Cin * Cout
            key = f"conv_wgrad_tcg[{Cin}->{Cout}]"
        elif n == "conv_wgrad_tc":
            B, Wp = args[3], args[5]
            fl = 2.0 * B * (Wp - 2) * (Wp - 2) * 9 * 32 * 32
        elif n == "conv_wgrad":
            B, Hs, Ws, Cin, Cout, pad, up = args[4], args[5], args[6], args[7], args[8], args[9], args[10]
            Ho = Hs * up + 2 * pad - 2
            fl = 2.0 * B * Ho * Ho * 9 * Cin * (9 if Cout == 16 else Cout)
        tot, cnt, flops, nbytes = fam.get(key, (0.0, 0, 0.0, 0.0))
        fam[key] = (tot + t, cnt + 1, flops + fl, nbytes + by)
        if n == "conv_chain":
            t0, c0, f0, b0 = fam.get("_conv3x3_tc_kernel", (0.0, 0, 0.0, 0.0))
            fam["_conv3x3_tc_kernel"] = (t0 + t, c0 + 1, f0 + fl, b0 + by)
            if args[1][0][4] == 0 and args[1][0][0] == 2 * PER_GPU_BATCH:      # forward chain over 256 samples: see ncu_traffic()
                t0, c0 = fam.get("_conv_tc_fwd_l1", (0.0, 0, 0.0, 0.0))[:2]
                fam["_conv_tc_fwd_l1"] = (t0 + t, c0 + 1, fl, by)
        if n == "conv_tc":                                          # the kernel behind both conv_tc families
            t0, c0, f0, b0 = fam.get("_conv3x3_tc_kernel", (0.0, 0, 0.0, 0.0))
            fam["_conv3x3_tc_kernel"] = (t0 + t, c0 + 1, f0 + fl, b0 + by)
            if args[11] == 0 and args[7] == 43 and args[6] == 2 * PER_GPU_BATCH:   # forward, 41x41 -> 39x39, 256 samples: see ncu_traffic()
                t0, c0 = fam.get("_conv_tc_fwd_l1", (0.0, 0, 0.0, 0.0))[:2]
                fam["_conv_tc_fwd_l1"] = (t0 + t, c0 + 1, fl, by)
    fam["_ingraph"] = ingraph
    return fam, nsteps


def time_torch_eager_gpu(B, steps=6, warmup=3):
    """The reference's arithmetic as PyTorch eager ON THE B200 (BASELINE.md 3 'second baseline'): the oracle port with its
    tensors on cuda:0, cuDNN convs with allow_tf32=True and fp32 cuBLAS matmuls (the reference's defaults), CUDA-event timed.
    Same synthetic inputs as config 1 / 2.  This is a baseline leg: none of the repo's kernels run here."""
    from oracle import sgsac_oracle as O
    from oracle.pin_rnd import make_rnd
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    args = O.Args(algorithm="sgsac", sgqn_quantile=0.95, batch_size=B)
    rs = np.random.RandomState(0)
    rep = O.synthetic_replay(1000, 2, seed=0)
    orc = O.make_oracle((9, 84, 84), (2,), args, seed=0)
    orc.p = type(orc.p)((k, v.to(dev)) for k, v in orc.p.items())
    orc.pool = torch.as_tensor(rs.randint(0, 256, size=(256, 3, 84, 84), dtype=np.uint8)).to(dev)
    L = NullLog()

    def one(step):
        batch = tuple(t.to(dev, non_blocking=True) for t in rep.sample(rs.randint(0, 1000, size=B)))
        rnd = make_rnd(rs, B, 2, 256)
        rnd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in rnd.items()}
        orc.update_from_batch(batch, rnd, L, step)

    step = 1
    for _ in range(warmup):
        one(step); step += 1
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        one(step); step += 1
    torch.cuda.synchronize()
    dt = (time.time() - t0) / steps
    return {"value": 1.0 / dt, "unit": "updates/s", "ms_per_step": dt * 1e3, "batch": B,
            "what": "oracle port of SGSAC.update as PyTorch eager on cuda:0 (cuDNN conv allow_tf32=True, cuBLAS fp32 matmul, "
                    "torch.sort-based quantile masks, per-tensor Adam), host-side sampling + H2D included, wall-clock over "
                    f"{steps} updates ({steps // 2} even + {steps - steps // 2} odd) after {warmup} warm-up"}


def build_info():
    """sgqn-carla_b200/BUILD_INFO.json (written by __graft_entry__.build()): which build of libsgqn_b200.so this process loaded."""
    p = os.path.join(ROOT, "sgqn-carla_b200", "BUILD_INFO.json")
    try:
        d = json.load(open(p))
        d["so_mtime"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(os.path.getmtime(os.path.join(ROOT, "sgqn-carla_b200", "libsgqn_b200.so"))))
        return d
    except Exception:
        return None


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed summary of the
    `ncu --set full` capture of THIS build (profiles/traffic_r2.json, written by profiles/summarize.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


def run_b200(a):
    import torch.distributed as dist
    import sgqn_carla_b200 as S
    from sgqn_carla_b200 import _lib
    from sgqn_carla_b200.dist import GradSync, P2PGradSync

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    sync = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        # gradient exchange: our peer-memory kernels (csrc/p2p.cu) unless SGQN_P2P=0 asks for the NCCL communicators
        sync = P2PGradSync() if os.environ.get("SGQN_P2P", "1") == "1" else GradSync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(algorithm="sgsac", B=PER_GPU_BATCH, A=2, size=84, storage="device", Bg=None, capacity=CAPACITY, data=None):
        """Agent + replay ring of one rank.  Bg = global batch (losses are means over it; None: single process)."""
        args = S.default_args(algorithm=algorithm, batch_size=B, sgqn_quantile=0.95, seed=1 + rank)
        ag = S.make_agent((9, size, size), (A,), args, dist=sync if Bg else None, global_batch=Bg)
        ag.engine.seed, ag.engine.seed_shared = 1234 + rank, 1234      # own indices / noise per rank, ONE fill scalar per global batch
        if Bg and world > 1:                # identical replicated parameters, targets, optimiser states
            ag.sync_from_rank0()
        frames, actions, rewards, not_dones, pool = data
        if algorithm == "sgsac":
            ag.set_overlay_pool(pool)
        if algorithm == "svea":
            ag.set_places_pool(torch.as_tensor(pool[:512]).float() / 255.0)
        rb = S.ReplayBuffer((9, size, size), (A,), capacity, B, storage=storage, frame_capacity=capacity + 8)
        rb.load_ring(frames, actions, rewards, not_dones)
        return ag, rb

    def timed(agent, rb, L, steps, warmup, s0=1, collective=True):
        step = s0
        for _ in range(warmup):
            agent.update(rb, L, step); step += 1
        barrier() if collective else torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.launch_count
        e0.record()
        for _ in range(steps):
            agent.update(rb, L, step); step += 1
        e1.record()
        barrier() if collective else torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1 and collective:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
        return ms, _lib.launch_count - c0

    def roofline_of(fam, nst, batch):
        """Roofline record of conv3x3_tc_kernel (the dominant kernel of every configuration) from a profile_kernels() pass."""
        tot = sum(v[0] for k, v in fam.items() if not k.startswith("_"))
        hbm, tf_burst, tf_sus, how = peaks()
        kt = fam["_conv3x3_tc_kernel"]                       # (ms, launches, flops, algorithmic bytes) over nst updates
        iso = kt[3] / (kt[0] * 1e-3) / 1e9
        ig = fam.get("_ingraph")
        name = "conv3x3_tc_kernel (32->32 SharedCNN layers, forward + data-gradient launches)"
        r = {"bound": "hbm", "kernel": name, "achieved": iso, "peak": hbm, "unit": "GB/s", "frac": iso / hbm, "traffic": None,
             "share_of_step": kt[0] / tot, "peak_source": f"HBM copy bandwidth {hbm} GB/s ({how})",
             "launches_per_step": kt[1] / nst, "ms_per_step_in_kernel": kt[0] / nst,
             "algorithmic_bytes_per_launch": kt[3] / kt[1], "avg_launch_us": kt[0] / kt[1] * 1e3,
             "tensor_view": {"achieved_tflops": kt[2] / (kt[0] * 1e-3) / 1e12, "peak_tflops": tf_sus / 2.0,
                             "peak_source": f"bf16 sustained {tf_sus} TF/s ({how}) / 2 = TF32 dense, derived"}}
        if ig is not None:
            # headline = the launches as the product issues them (inside the update's CUDA graph, back to back with programmatic
            # dependent launch); the isolated-launch figure (launch latency + cold pipeline + idle GPU before every launch) beside it
            a = ig["bytes_per_launch"] / (ig["us_per_launch"] * 1e-6) / 1e9
            r.update({"achieved": a, "frac": a / hbm, "avg_launch_us": ig["us_per_launch"],
                      "algorithmic_bytes_per_launch": ig["bytes_per_launch"], "ms_per_step_in_kernel": ig["us_per_launch"] * kt[1] / nst * 1e-3,
                      "tensor_view": dict(r["tensor_view"], achieved_tflops=ig["flops_per_launch"] / (ig["us_per_launch"] * 1e-6) / 1e12),
                      "timing": f"CUDA events around 20 replays of a CUDA graph that holds the {ig['launches']} conv3x3_tc_kernel launches of one odd + one even "
                                "update in program order on one stream (how the update's own graph runs them: programmatic dependent launch, "
                                "each layer's input still in L2); average per launch",
                      "isolated_launch": {"achieved": iso, "frac": iso / hbm, "avg_launch_us": kt[0] / kt[1] * 1e3,
                                          "timing": "CUDA events around every launch in an eager single-stream pass over 4 updates (idle GPU before each launch)"}})
        return r

    data84 = synthetic(CAPACITY, 2, seed=rank)
    B = PER_GPU_BATCH
    Bg = B * world

    # ---- device-resident run (value): weak scaling, 128 samples per rank
    agent, rb = make(Bg=Bg, data=data84)
    L = NullLog()
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(agent, rb, L, a.steps, a.warmup)
    sampler.stop_flag = True
    last = {k: float(v) for k, v in L.last.items()}
    assert all(np.isfinite(v) for v in last.values()), last
    ups = a.steps / (ms / 1000.0)
    value = ups * (Bg / PER_GPU_BATCH)
    sharded = None
    if world > 1:
        # the replicas must still hold bit-identical parameters after the timed updates (every rank applied the same summed
        # gradients): element-wise max == min over the ranks, outside the timed region
        torch.cuda.synchronize()
        mx, mn = agent.engine.params.clone(), agent.engine.params.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sharded = {"replicas_bit_identical": bool(torch.equal(mx, mn)), "parameters_compared": int(mx.numel()),
                   "updates_before_check": a.steps + a.warmup}
        if getattr(sync, "arena", None) is not None:
            sharded["p2p_barrier_timeouts"] = sync.timeouts()
        del mx, mn

    # ---- kernel-family profile + roofline of the dominant family (rank 0)
    roof, fam_rows = None, None
    fam, nst = profile_kernels(agent, rb, 4)          # every rank runs it (the updates contain collectives)
    if rank == 0:
        tot = sum(v[0] for k, v in fam.items() if not k.startswith("_"))
        fam_rows = sorted(((k, v[0] / nst, v[1] // nst, v[2] / nst) for k, v in fam.items() if not k.startswith("_")), key=lambda r: -r[1])
        # conv3x3_tc_kernel (SharedCNN 32->32 layers, forward + data gradient) is the step's dominant kernel.  72 (fwd) /
        # 48 (dgrad) FLOP per algorithmic byte is below the B200 ridge (TF32 692 TF/s / 6.5 TB/s = 106 FLOP/B): HBM-bound.
        roof = roofline_of(fam, nst, B)
        roof.setdefault("timing", "CUDA events around every launch of the kernel in an eager (graph-free, single-stream) pass over 4 updates, the "
                        "stream parked behind a spin kernel while the host queues each update so the events bracket device time only")
        l1 = fam.get("_conv_tc_fwd_l1")
        tr = ncu_traffic()
        if l1 is not None and tr is not None:
            # one specific launch, so that `traffic` (ncu --set full of this build, profiles/) and the live time refer to the
            # same work: forward 41x41 -> 39x39 at B = 128 (2B = 256 samples in the [next_obs ; obs] pass)
            us = l1[0] / l1[1] * 1e3
            roof.update({"traffic": tr.get("dram_bytes_per_launch"),
                         "traffic_launch": {"what": tr.get("what"), "source": tr.get("source"), "algorithmic_bytes": tr.get("algorithmic_bytes"),
                                            "live_us_same_shape": us, "ncu_us_cold_cache": tr.get("duration_us")}})
    if world > 1:
        barrier()
    del rb

    # ---- BASELINE config 4 proper: global batch 1024 sharded over the ranks (strong scaling; N=1: all 1024 on one GPU)
    strong = None
    if not a.quick and 1024 % world == 0:
        Bs = 1024 // world
        ag4, rb4 = make(B=Bs, Bg=1024, data=data84)
        st4 = max(20, a.steps // 4)
        ms4, _ = timed(ag4, rb4, NullLog(), st4, max(3, a.warmup // 2))
        strong = {"global_batch": 1024, "per_gpu_batch": Bs, "ms_per_step": ms4 / st4, "updates_per_s": st4 / (ms4 / 1e3),
                  "batch128_equiv_updates_per_s": st4 / (ms4 / 1e3) * 8.0, "steps": st4, "scaling": "strong",
                  "what": "BASELINE config 4: SGSAC, global batch 1024 split evenly over the ranks, gradients summed across the ranks ("
                          + ("own kernels over NVLink peer memory" if getattr(sync, "arena", None) is not None else
                             ("NCCL" if sync is not None else "single rank: no exchange")) + ")"}
        del ag4, rb4

    # ---- end-to-end run: host-resident replay ring (pinned), loss read-back every step
    agent2, rb2 = make(storage="pinned", Bg=Bg, data=data84)
    agent2.defer_logs = True
    L2 = NullLog()
    ms2, _ = timed(agent2, rb2, L2, a.steps, a.warmup)
    _ = {k: float(v) for k, v in L2.last.items()}
    e2e_v = a.steps / (ms2 / 1000.0) * (Bg / PER_GPU_BATCH)
    h2d = 2 * B * 9 * 84 * 84 + B * 8 * 2 + 8                 # sampled uint8 stacks pulled from pinned host memory (+ nothing else: actions etc. live on device)
    d2h = 8 * 4

    # ---- acting latency: host uint8 stack -> action on the host, one CUDA graph launch per call
    act = None
    if rank == 0:
        ob = data84[0][:3].reshape(9, 84, 84)
        for fn_name in ("select_action", "sample_action"):
            fn = getattr(agent2, fn_name)
            for _ in range(5):
                fn(ob)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter(); fn(ob); ts.append(time.perf_counter() - t0)
            act = dict(act or {}, **{fn_name + "_us": float(np.median(ts) * 1e6)})
        act["how"] = "median host wall time of 200 calls, uint8 (9,84,84) host array in -> float32 (A,) host array out"
    del agent2, rb2
    if world > 1:
        barrier()

    if rank != 0:
        _finish(world)
        return

    # ---- the other single-GPU configurations of BASELINE.json (rank 0, no collectives): config 3 and config 5
    others = None
    if world == 1 and not a.quick:
        others = {}
        st = max(40, a.steps // 2)
        specs = [("config3_sgsac_b256", "sgsac", 256, 2, 84, "BASELINE config 3: SGSAC on CARLA-shaped observations (9x84x84 uint8, A=2), batch 256, overlay pool 2048"),
                 ("config5_svea_b128", "svea", 128, 6, 84, "BASELINE config 5: SVEA (random_shift pad 4, critic on [obs ; overlay(obs)] = 256 rows), A=6, batch 128"),
                 ("config5_rad_b128", "rad", 128, 6, 100, "BASELINE config 5: RAD (100x100 frames, random_crop to 84 fused into the gather), A=6, batch 128")]
        for key, algo, Bc, Ac, size, what in specs:
            d = data84 if (Ac == 2 and size == 84) else synthetic_sized(8000, Ac, size, seed=3)
            agc, rbc = make(algorithm=algo, B=Bc, A=Ac, size=size, data=d, capacity=len(d[1]))
            msc, lc = timed(agc, rbc, NullLog(), st, max(3, a.warmup), collective=False)
            famc, nc = profile_kernels(agc, rbc, 4)
            r = roofline_of(famc, nc, Bc)
            others[key] = {"what": what, "updates_per_s": st / (msc / 1e3), "ms_per_step": msc / st, "steps": st, "gpu_launches": lc,
                           "batch128_equiv_updates_per_s": st / (msc / 1e3) * Bc / 128.0,
                           "roofline": {k: r[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "share_of_step",
                                                          "launches_per_step", "ms_per_step_in_kernel", "avg_launch_us")}}
            del agc, rbc
    cpu, eager = None, None
    if world == 1 and not a.no_cpu_baseline:
        v, cores, sample = time_oracle(4, 1, 40.0)
        cpu = {"value": v, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample}
        if not a.quick:
            try:
                eager = {"b128": time_torch_eager_gpu(128), "b256": time_torch_eager_gpu(256, steps=4, warmup=2)}
            except Exception as e:          # a baseline leg must never take the measurement down
                eager = {"error": repr(e)[:300]}
    gflop = 0.5 * (GFLOP_ODD + GFLOP_EVEN) * Bg
    line = {
        "metric": "SGSAC updates/sec (batch 128, 9x84x84)", "value": value, "unit": "updates/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "per_gpu_batch": B, "global_batch": Bg, "parallelism": f"dp{world}" if world > 1 else "single",
                   "replay_capacity": CAPACITY, "l2_policy": "inputs larger than L2 (423 MB frame ring, random gather; ~390 MB activations per encoder pass)",
                   "value_definition": "global updates/s x (global_batch/128)",
                   "cuda_graphs": bool(agent.use_cuda_graphs), "conv_precision": agent.engine.precision,
                   "collectives": "none" if sync is None else ("own kernels over NVLink peer memory (csrc/p2p.cu)"
                                                               if getattr(sync, "arena", None) is not None else "NCCL")},
        "e2e": {"value": e2e_v, "unit": "updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "agent.update(replay_buffer, L, step) with the replay frame ring in pinned HOST memory: every step the sampled "
                       "uint8 frames of the NEXT batch cross PCIe into a device staging ring (zero-copy kernel on a side stream, under "
                       "the current update), are converted / cropped on the device at the start of their step, and the step's loss "
                       "vector is copied device->host"},
        "act_latency": act, "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
        "torch_eager_b200": eager, "config4_strong": strong, "sharded_check": sharded, "other_configs": others,
        "algorithmic_gflop_per_update": gflop, "achieved_tflops_whole_step": gflop * ups / 1e3,
        "kernel_families_ms_per_step": [[r[0], round(r[1], 4), r[2]] for r in (fam_rows or [])[:12]],
        "losses_last_step": last, "build_info": build_info(),
    }
    emit(line)
    _finish(world)


def _finish(world):
    """Multi-rank exit: captured CUDA graphs that contain NCCL kernels can deadlock destroy_process_group(); leave together
    after a barrier and let process exit tear the communicators down."""
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


class _StdoutToStderr(object):
    """Everything any library prints to fd 1 while the benchmark runs (NCCL's version banner, ...) goes to stderr, so
    stdout carries exactly the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        sys.stdout.flush()
        os.dup2(self.saved, 1)

    def __exit__(self, *a):
        self.restore()
        return False


_REDIRECT = None


def emit(line):
    if _REDIRECT is not None:
        _REDIRECT.restore()
    print(json.dumps(line), flush=True)


def main():
    global _REDIRECT
    _REDIRECT = _StdoutToStderr().__enter__()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline numbers only (skip config 3 / 4-strong / 5 and the torch-eager baseline)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
