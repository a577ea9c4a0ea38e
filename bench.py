#!/usr/bin/env python
"""bench.py -- PBS/s of the batched programmable bootstrap (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--preset P1] [--batch 4096] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

A "step" is one pass of the hot path (blind rotation + sample extract + key switch) over one batch of
synthetic LWE ciphertexts.  Headline workload = BASELINE.json configs[1]: batch 4096 PBS per GPU, N=1024,
n=630, identity test vector (preset P1 of SURVEY.md 8(d)).

value  : N = 1: inputs resident in HBM (CUDA tensors handed to the C ABI as device pointers).
         N > 1 (torchrun, one rank per GPU): the WHOLE batch of N x 4096 ciphertexts lives on rank 0; a step is
         NCCL scatter -> every rank bootstraps its shard with its replica of the keys -> NCCL gather back to rank 0
         (sharding.bootstrap_sharded), all inside the timed region; weak scaling (fixed work per GPU).
         `replicas` carries the same step without the scatter/gather for comparison.
e2e    : the same through pinned HOST buffers: H2D (N > 1: on rank 0, then the scatter) and D2H inside the timed region.
configs: one sub-record per other BASELINE config -- P0 (the reference's default parameters) batch 4096, P2 batch
         16384 with a programmable LUT, the BMMP variant, the single bootstrapped NAND of config #1 with its one-core
         CPU figure, and the sharded legs of configs #4 / #5: a 65 536-ciphertext batch (strong scaling), the 65 536-gate
         circuit (depth-1 and 16 x 4096 layered, one all-gather per level) and the BMMP batch of 65 536.
roofline: integer pipe (IMAD-class lane-ops, SURVEY 8(d) W_int) against the IMAD peak measured live
         on this GPU by tfhe_measure_int_peak; the pipes the kernel really runs on are in roofline.fp64 / roofline.smem.
cpu_baseline / --impl reference: the C restatement of the reference's Rust path (oracle/, faithful
         Toeplitz O(N^2) algorithm, one PBS per host thread) -- the Rust crate cannot be built here.  The reference arm
         imports nothing but oracle/.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md 8(d) presets (the b200 arm asserts they equal tfhe_params_preset; the reference arm feeds them to the oracle)
_P0 = dict(glwe_dimension=2, glwe_poly_degree=9, lwe_dimension=722, padding_bits=1, log_p=2, log_q=32, ks_log_base=4, ks_levels=5,
           pbs_log_base=4, pbs_levels=6, lwe_std_dev=0.000013071021089943935, glwe_std_dev=0.00000004990272175010415)
PRESETS = {
    "P0": _P0,
    "P1": dict(_P0, glwe_dimension=1, glwe_poly_degree=10, lwe_dimension=630, pbs_log_base=8, pbs_levels=3, ks_log_base=2, ks_levels=8,
               log_p=2, lwe_std_dev=3.0517578125e-05, glwe_std_dev=2.9802322387695312e-08),
    "P2": dict(_P0, glwe_dimension=1, glwe_poly_degree=11, lwe_dimension=742, pbs_log_base=8, pbs_levels=3, ks_log_base=4, ks_levels=5,
               log_p=4, lwe_std_dev=7.069849454709433e-06, glwe_std_dev=4.656612873077393e-10),
}
KEY_SEED, INPUT_SEED = 0xB200, 1


def pview(fields):
    f = SimpleNamespace(**fields)
    f.k, f.N, f.n = f.glwe_dimension, 1 << f.glwe_poly_degree, f.lwe_dimension
    return f


def workload_string(preset, batch, f):
    return (f"{preset}: batch {batch} PBS per GPU, k={f.k} N={f.N} n={f.n} pbs(logB={f.pbs_log_base},l={f.pbs_levels}) "
            f"ks(logB={f.ks_log_base},l={f.ks_levels}) log_p={f.log_p}, identity test vector")


def config_dict(preset, batch, f):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": workload_string(preset, batch, f), "preset": preset, "batch_per_gpu": batch,
            "keys": f"seeded keygen (seed {KEY_SEED:#x}), replicated per GPU", "inputs": f"seeded encryptions (seed {INPUT_SEED}) of messages i mod 2^log_p",
            "l2": "256 MB flush write between timed iterations (GPU arm)"}


def w_int_per_pbs(p):
    """IMAD-class multiply instructions per PBS (algorithmic, kernel independent)."""
    P, l, N, R = p.k + 1, p.pbs_levels, p.N, 2
    bf = (N // 2) * p.glwe_poly_degree
    c_cmux = 3 * R * bf * (P * l + P) + R * P * P * l * N + 8 * P * N
    c_ks = p.k * N * p.ks_levels * (p.n + 1)
    return p.n * c_cmux + c_ks, c_cmux, c_ks


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every 20 ms through NVML from a thread of this process (the
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` fields); falls back to an nvidia-smi
    subprocess.  start() before the warm-up, mark() at the start of the timed region: samples after the mark are
    reported (the warm-up runs the same kernels, so for a very short timed region the last warm-up samples stand in)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx, self.samples, self.stop_flag, self.t, self.mark_at, self.how = gpu_index, [], False, None, 0, None
        self.sm_max = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.idx)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.idx)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)

            def loop():
                while not self.stop_flag:
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    except Exception:
                        r = 0
                    self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(r)))
                    time.sleep(0.02)
            self.how = "nvml"
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception:
            self.how = "nvidia-smi"
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

                def pump():
                    for ln in self.proc.stdout:
                        f = [x.strip() for x in ln.split(",")]
                        try:
                            bits = sum(m for (_, m), v in zip(self.REASONS, f[2:6]) if v.lower().startswith("active"))
                            self.samples.append((float(f[0]), bits))
                            self.sm_max = float(f[1])
                        except (ValueError, IndexError):
                            pass
                self.t = threading.Thread(target=pump, daemon=True)
                self.t.start()
            except Exception:
                self.how = None

    def mark(self):
        self.mark_at = len(self.samples)

    def stop(self):
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.stop_flag = True
        if self.how == "nvidia-smi":
            self.proc.terminate()
        self.t.join(timeout=2)
        timed = self.samples[self.mark_at:]
        use, src = (timed, "timed region") if len(timed) >= 2 else (self.samples[-8:], "end of warm-up + timed region")
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
        reasons = sorted({name for _, bits in use for name, m in self.REASONS if bits & m})
        return {"sm_mhz": float(np.median([c for c, _ in use])), "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(use),
                "sampled": f"{self.how}, {src}"}


# ------------------------------------------------------------------------------------------------ CPU (oracle) legs
def cpu_reference_rate(o, n_ct, threads, bsk, ksk, cts, tv):
    """PBS/s of the oracle (faithful reference algorithm: N x N Toeplitz matrix per product), one PBS per host thread."""
    from oracle import orc
    orc.lib().orc_set_faithful_toeplitz(1)
    t0 = time.perf_counter()
    out = orc.bootstrap_batch(o, cts[:n_ct], bsk, ksk, tv, threads)
    dt = time.perf_counter() - t0
    orc.lib().orc_set_faithful_toeplitz(0)
    return n_ct / dt, dt, out


def reference_arm(args, fields, cores):
    """--impl reference: the reference's CPU path (C port in oracle/), all host threads; imports oracle/ only."""
    from oracle import orc
    f = pview(fields)
    o = orc.params(**fields)
    lwe_sk, glwe_sk, bsk, ksk = orc.keygen(o, KEY_SEED)
    tv = orc.test_vector_identity(o)
    pm = 1 << f.log_p
    n_ct = cores
    cts = np.stack([orc.lwe_encrypt(o, lwe_sk, i % pm, INPUT_SEED, i) for i in range(n_ct)])
    times = []
    for s in range(args.warmup + args.steps):
        rate, dt, out = cpu_reference_rate(o, n_ct, cores, bsk, ksk, cts, tv)
        if s >= args.warmup:
            times.append(dt)
        if s == 0:
            assert all(orc.lwe_decrypt_round(o, lwe_sk, out[i]) == i % pm for i in range(n_ct))
    total = sum(times)
    value = n_ct * len(times) / total
    sample = (f"{n_ct} PBS per step (one per host thread) of the same workload, C restatement of the reference's Rust path, "
              "Toeplitz O(N^2) algorithm (utils.rs:113-160)")
    return {"impl": "reference", "metric": "PBS/sec", "value": value, "unit": "PBS/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config_dict(args.preset, args.batch, f),
            "cpu_baseline": {"value": value, "unit": "PBS/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------ GPU arm
class Env:
    """One parameter set on this rank's GPU: keys (seeded: identical on every rank), context on `stream`, device key."""

    def __init__(self, T, preset, device, stream, bmmp=False):
        self.T, self.preset, self.bmmp = T, preset, bmmp
        self.p = T.TfheParams.preset(preset)
        for name, v in PRESETS[preset].items():
            assert getattr(self.p, name) == v, f"bench.py preset table out of date: {preset}.{name}"
        self.ctx = T.Context(self.p, device)
        self.ctx.set_stream(stream.cuda_stream)   # torch CUDA events on `stream` bracket exactly the kernels of this ctx
        if bmmp:
            self.lwe_sk, self.glwe_sk, self.bsk, self.ksk = T.bootstrapping_key_gen_bmmp(self.p, KEY_SEED)
            self.bk = self.ctx.upload_key_bmmp(self.bsk, self.ksk)
        else:
            self.lwe_sk, self.glwe_sk, self.bsk, self.ksk = T.bootstrapping_key_gen(self.p, KEY_SEED)
            self.bk = self.ctx.upload_key(self.bsk, self.ksk)
        self.pm = 1 << self.p.log_p
        self.tv = T.construct_identity_test_vector(self.p)

    def enc(self, m, idx):
        return self.T.encrypt_lwe_plaintext(self.p, self.lwe_sk, self.T.encode_message(self.p, m), INPUT_SEED, idx)

    def dec(self, ct):
        return self.T.decode_rounded(self.p, self.T.decrypt_lwe(self.lwe_sk, ct))

    def batch(self, B, base=0, n_unique=256):
        """B ciphertexts of messages (base + i) mod 2^log_p: `n_unique` real encryptions, tiled (timing is data independent)."""
        nu = min(B, n_unique)
        uniq = np.stack([self.enc((base + i) % self.pm, base + i) for i in range(nu)])
        return np.tile(uniq, ((B + nu - 1) // nu, 1))[:B].copy(), nu

    def close(self):
        self.bk.free()
        self.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--preset", default="P1")
    ap.add_argument("--batch", type=int, default=4096, help="ciphertexts per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the sub-records of the other BASELINE configs")
    ap.add_argument("--cpu-threads", type=int, default=0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = args.cpu_threads or (os.cpu_count() or 1)
    fields = PRESETS[args.preset]

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, fields, cores)))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PBS path has no CPU fallback (use --impl reference for the CPU arm)")
    import tfhe_research_b200 as T
    from tfhe_research_b200 import circuit, sharding
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits that keep the idle ranks' GPUs free (mgpu leg)
    stream = torch.cuda.Stream()        # a real (non-default) stream shared by torch events, NCCL ordering and the C-ABI contexts
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    i32 = torch.int32

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*xs):
        if world == 1:
            return list(xs)
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def timed(fn, nsteps, after=None):
        """nsteps x (L2 flush, then fn bracketed by CUDA events on the stream); barrier + synchronise on both sides."""
        barrier()
        evs, extra = [], []
        for _ in range(nsteps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            evs.append((e0, e1))
            if after:
                extra.append(after())
        barrier()
        return [a.elapsed_time(b) for a, b in evs], extra

    def dev(x_np):
        return torch.from_numpy(np.ascontiguousarray(x_np).view(np.int32)).cuda()

    def sharded_step(env, d_tv, root_in, total, out_root):
        """NCCL scatter from rank 0 -> local bootstrap -> NCCL gather into out_root on rank 0 (world 1: just the bootstrap)."""
        return sharding.bootstrap_sharded(lambda x: env.ctx.bootstrap(env.bk, x, d_tv), root_in, total, env.p.n + 1, "cuda", i32, out=out_root)

    # ================================================================ headline: BASELINE configs[1]
    E = Env(T, args.preset, local_rank, stream)
    p, ctx, bk = E.p, E.ctx, E.bk
    w_int, c_cmux, c_ks = w_int_per_pbs(p)
    B = args.batch
    total = world * B
    row = p.n + 1
    cts_local, n_unique = E.batch(B, base=rank * B)
    host_tv = torch.from_numpy(E.tv.view(np.int32).copy()).pin_memory()
    d_tv = host_tv.cuda()
    d_in = dev(cts_local)
    d_out = torch.empty((B, row), dtype=i32, device="cuda")
    if world > 1:
        # the whole job's batch lives on rank 0 (rank r's shard = the ciphertexts rank r would encrypt itself)
        if rank == 0:
            full = np.concatenate([E.batch(B, base=r * B)[0] for r in range(world)])
            host_root_in = torch.from_numpy(full.view(np.int32).copy()).pin_memory()
            host_root_out = torch.empty((total, row), dtype=i32).pin_memory()
            root_in = host_root_in.cuda()
            root_out = torch.empty((total, row), dtype=i32, device="cuda")
        else:
            host_root_in = host_root_out = root_in = root_out = None
    host_in = torch.from_numpy(cts_local.view(np.int32).copy()).pin_memory()
    host_out = torch.empty((B, row), dtype=i32).pin_memory()

    peaks = ctx.measure_int_peak()
    fft_path = ctx.pbs_path == T.PATH_FFT
    fp64_peaks = ctx.measure_fp64_peak() if fft_path else None

    def step_replica():
        ctx.bootstrap(bk, d_in, d_tv, out=d_out)

    def step_sharded():
        sharded_step(E, d_tv, root_in, total, root_out)

    def step_e2e():
        if world == 1:
            ctx.bootstrap(bk, host_in, host_tv, out=host_out)     # the C ABI copies in and out on the stream
        else:
            rin = host_root_in.cuda(non_blocking=True) if rank == 0 else None
            res = sharded_step(E, d_tv, rin, total, root_out)
            if rank == 0:
                host_root_out.copy_(res, non_blocking=True)

    headline = step_replica if world == 1 else step_sharded
    sampler = ClockSampler(local_rank)
    sampler.start()
    timed(headline, args.warmup)
    launches0 = ctx.launch_count
    sampler.mark()
    step_ms, kern = timed(headline, args.steps, after=ctx.last_timing)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    timed(step_e2e, 1)
    e2e_ms, _ = timed(step_e2e, args.steps)
    rep_ms = step_ms
    if world > 1:
        timed(step_replica, 1)
        rep_ms, _ = timed(step_replica, args.steps)

    # per-PBS latency: one ciphertext through the same call (n dependent CMUX steps + key switch)
    def latency_batch1(env, d_tv_, one_in):
        one_out = torch.empty((1, env.p.n + 1), dtype=i32, device="cuda")
        lat = []
        for _ in range(4):
            env.ctx.bootstrap(env.bk, one_in, d_tv_, out=one_out)
            lat.append(env.ctx.last_timing()["total_ms"])
        return min(lat[1:])
    latency_batch1_ms = latency_batch1(E, d_tv, d_in[:1].contiguous())

    # correctness of what was timed: decrypt a sample of the last outputs
    res = d_out.cpu().numpy().view(np.uint32)
    if world == 1:
        assert np.array_equal(res, host_out.numpy().view(np.uint32)), "device-pointer and host-pointer paths disagree"
    for i in range(0, B, max(1, B // 64)):
        assert E.dec(res[i]) == (rank * B + i % n_unique) % E.pm, f"PBS {i} decrypts wrongly"
    if world > 1 and rank == 0:
        gathered = root_out.cpu().numpy().view(np.uint32)
        assert np.array_equal(gathered, host_root_out.numpy().view(np.uint32)), "device and host sharded paths disagree"
        assert np.array_equal(gathered[:B], res), "rank 0's shard of the gathered result differs from its local result"
        for i in range(0, total, max(1, total // 128)):
            assert E.dec(gathered[i]) == ((i // B) * B + (i % B) % n_unique) % E.pm, f"gathered PBS {i} decrypts wrongly"

    t_dev, t_e2e, t_rep = sum(step_ms), sum(e2e_ms), sum(rep_ms)
    t_br = sum(k["blind_rotate_ms"] for k in kern) / len(kern)
    t_ks = sum(k["key_switch_ms"] for k in kern) / len(kern)
    t_dev, t_e2e, t_rep, t_br = max_over_ranks(t_dev, t_e2e, t_rep, t_br)

    def frac_of(pp, batch, br_ms):
        return batch * pp.n * w_int_per_pbs(pp)[1] / (br_ms * 1e-3) / peaks["imad"]

    # ================================================================ the other BASELINE configs (sub-records)
    configs = {}

    def single_gpu_config(name, preset, batch, bmmp=False, lut="identity", steps=2, warmup=1):
        """One more parameter set / variant on this GPU: device-resident throughput, kernel time, roofline fraction, latency."""
        env = Env(T, preset, local_rank, stream, bmmp=bmmp)
        cts, nu = env.batch(batch)
        x, out = dev(cts), torch.empty((batch, env.p.n + 1), dtype=i32, device="cuda")
        if lut == "programmable":        # config #3: a random f: Z16 -> Z16 with f(0) = 0, seed 2
            f = np.random.default_rng(2).integers(0, env.pm, env.pm).astype(np.uint32)
            f[0] = 0
            tv = T.construct_test_from_lut(env.p, f)
        else:
            f, tv = np.arange(env.pm, dtype=np.uint32), env.tv
        d_tv_ = dev(tv)
        timed(lambda: env.ctx.bootstrap(env.bk, x, d_tv_, out=out), warmup)
        ms, kt = timed(lambda: env.ctx.bootstrap(env.bk, x, d_tv_, out=out), steps, after=env.ctx.last_timing)
        r = out.cpu().numpy().view(np.uint32)
        for i in range(0, batch, max(1, batch // 32)):
            assert env.dec(r[i]) == int(f[(i % nu) % env.pm]), f"{name}: PBS {i} decrypts wrongly"
        br = sum(k["blind_rotate_ms"] for k in kt) / len(kt)
        rec = {"workload": workload_string(preset, batch, env.p).replace("identity test vector", f"{lut} test vector") + (", BMMP unrolled-by-two key triples" if bmmp else ""),
               "value": batch * steps / (sum(ms) * 1e-3), "unit": "PBS/s", "ms_per_step": sum(ms) / steps, "steps": steps, "warmup": warmup,
               "kernel_ms": br, "key_switch_ms": sum(k["key_switch_ms"] for k in kt) / len(kt),
               "roofline": {"frac": frac_of(env.p, batch, br), "bound": "int32-imad",
                            "note": "SURVEY 8(d) W_int of the standard n-step chain over the measured IMAD peak" + (" (the BMMP kernel runs n/2 steps with 3 keys each)" if bmmp else "")},
               "latency_ms_batch1": latency_batch1(env, d_tv_, x[:1].contiguous())}
        return rec, env

    if world == 1 and not args.no_configs:
        rec, env0 = single_gpu_config("P0_b4096", "P0", 4096)
        configs["P0_b4096"] = rec
        # ---- config #1: ONE bootstrapped NAND with the reference's default parameters, GPU latency and one-core CPU figure
        c0, c1 = env0.enc(1, 9001)[None], env0.enc(1, 9002)[None]
        d0, d1 = dev(c0), dev(c1)
        lat = []
        for _ in range(5):
            g = env0.ctx.gate(env0.bk, T.NAND, d0, d1)
            lat.append(env0.ctx.last_timing()["total_ms"])
        gate_out = g.cpu().numpy().view(np.uint32)[0]
        assert env0.dec(gate_out) == 0
        cfg1 = {"workload": "one bootstrapped NAND (trivial(1) - AND, SURVEY 9-B H6), reference default parameters (lib.rs:101-123)",
                "gpu_ms": min(lat[1:]), "gpu_ms_all": lat[1:]}
        if not args.no_cpu_baseline:
            from oracle import orc
            o0 = orc.params(**PRESETS["P0"])
            orc.lib().orc_set_faithful_toeplitz(1)
            t0 = time.perf_counter()
            cpu_gate = orc.gate(o0, 3, c0[0], c1[0], env0.bsk, env0.ksk)
            cfg1["cpu_ms_one_core"] = 1e3 * (time.perf_counter() - t0)
            orc.lib().orc_set_faithful_toeplitz(0)
            assert np.array_equal(cpu_gate, gate_out), "GPU NAND differs from the CPU oracle"
            cfg1["cpu_kind"] = "port (C restatement of boolean.rs `and` + bootstrap, Toeplitz O(N^2)); GPU output bit-identical"
        configs["cfg1_single_nand"] = cfg1
        env0.close()
        rec, env2 = single_gpu_config("P2_b16384", "P2", 16384, lut="programmable")
        configs["P2_b16384"] = rec
        env2.close()
        rec, envb = single_gpu_config("P1_bmmp_b4096", "P1", 4096, bmmp=True)
        configs["P1_bmmp_b4096"] = rec
        envb.close()

    if not args.no_configs:
        # ---- strong scaling: ONE batch of 65 536 ciphertexts on rank 0 -> scatter -> PBS -> gather (P1, and the BMMP variant = config #5)
        def strong_leg(env, total_cts, steps=2):
            d_tv_ = dev(env.tv)
            if rank == 0:
                cts, nu = env.batch(total_cts)
                rin, rout = dev(cts), torch.empty((total_cts, env.p.n + 1), dtype=i32, device="cuda")
            else:
                nu, rin, rout = min(total_cts, 256), None, None
            got = []
            timed(lambda: sharded_step(env, d_tv_, rin, total_cts, rout), 1)
            ms, _ = timed(lambda: got.append(sharded_step(env, d_tv_, rin, total_cts, rout)), steps)
            (t,) = max_over_ranks(sum(ms))
            if rank == 0:
                r = got[-1].cpu().numpy().view(np.uint32)
                for i in range(0, total_cts, total_cts // 64):
                    assert env.dec(r[i]) == (i % nu) % env.pm, f"sharded PBS {i} decrypts wrongly"
            return {"value": total_cts * steps / (t * 1e-3), "unit": "PBS/s", "total_ciphertexts": total_cts, "ms_per_step": t / steps, "steps": steps,
                    "scaling": "strong", "collectives": "NCCL send/recv scatter from rank 0 + gather to rank 0 inside the timed region" if world > 1 else "none (1 GPU)"}
        rec = strong_leg(E, 65536)
        rec["workload"] = "P1: ONE batch of 65536 PBS held on rank 0, sharded over all GPUs"
        configs["sharded_P1_b65536"] = rec
        envb = Env(T, "P1", local_rank, stream, bmmp=True)
        rec = strong_leg(envb, 65536)
        rec["workload"] = "config #5: BMMP unrolled-by-two bootstrapping + key switch, ONE batch of 65536 held on rank 0, sharded over all GPUs"
        if world == 1 and not args.no_cpu_baseline:
            from oracle import orc
            ob = orc.params(**PRESETS["P1"])
            cts, _ = envb.batch(cores)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:        # ctypes releases the GIL: one PBS per host thread
                outs = list(ex.map(lambda c: orc.bootstrap_bmmp(ob, c, envb.bsk, envb.ksk, envb.tv), cts))
            dt = time.perf_counter() - t0
            same = envb.ctx.bootstrap(envb.bk, cts, envb.tv)
            assert np.array_equal(same, np.stack(outs)), "GPU BMMP result differs from the CPU oracle"
            rec["cpu_baseline"] = {"value": cores / dt, "unit": "PBS/s", "cores": cores, "kind": "port",
                                   "sample": f"{cores} BMMP PBS, one per host thread, {dt:.1f} s; oracle restatement of notes/BMMP Bootstrapping.md (exact u32 bundle, "
                                             "direct O(N^2) products); GPU output bit-identical on this sample"}
        configs["bmmp_P1_b65536"] = rec
        envb.close()

        # ---- config #4: 65 536 NAND/AND/XOR gates on the reference's default parameters, sharded
        envc = Env(T, "P0", local_rank, stream)
        rng = np.random.default_rng(3)
        n_in = 1024
        bits = rng.integers(0, 2, n_in)
        wires_np = np.stack([envc.enc(int(b), i) for i, b in enumerate(bits)])     # every rank encrypts the same inputs (seeded)
        wires = dev(wires_np)
        crow = envc.p.n + 1

        def gate_fn(ops, ct0, ct1):
            return envc.ctx.gate(envc.bk, np.ascontiguousarray(ops), ct0, ct1)
        G = 65536
        ops = rng.choice(np.array([T.NAND, T.AND, T.XOR], dtype=np.uint8), G)
        il, ir = rng.integers(0, n_in, G), rng.integers(0, n_in, G)
        if rank == 0:
            root0 = wires.index_select(0, torch.as_tensor(ir, device="cuda"))      # right input -> ct0 (boolean.rs:18)
            root1 = wires.index_select(0, torch.as_tensor(il, device="cuda"))      # left  input -> ct1
            gout = torch.empty((G, crow), dtype=i32, device="cuda")
        else:
            root0 = root1 = gout = None
        lo, hi = sharding.shard_range(G, rank, world)

        def depth1():
            c0 = sharding.scatter_rows(root0, (crow,), G, "cuda", i32)
            c1 = sharding.scatter_rows(root1, (crow,), G, "cuda", i32)
            o = gate_fn(ops[lo:hi], c0, c1) if hi > lo else c0
            return sharding.gather_rows(o, G, out=gout)
        got1 = []
        timed(depth1, 1)
        ms1, _ = timed(lambda: got1.append(depth1()), 1)
        levels = circuit.random_layered_circuit(n_in, [4096] * 16, seed=3)
        timed(lambda: circuit.evaluate_encrypted(levels[:1], wires, gate_fn), 1)
        last = []
        ms2, _ = timed(lambda: last.append(circuit.evaluate_encrypted(levels, wires, gate_fn)), 1)
        t1, t2 = max_over_ranks(sum(ms1), sum(ms2))
        if rank == 0:
            fplain = {T.NAND: lambda l, r: 1 - (l & r), T.AND: lambda l, r: l & r, T.XOR: lambda l, r: l ^ r}
            g = got1[-1].cpu().numpy().view(np.uint32)
            for i in range(0, G, 257):
                assert envc.dec(g[i]) == fplain[int(ops[i])](int(bits[il[i]]), int(bits[ir[i]])), f"gate {i} decrypts wrongly"
            lw = last[-1].cpu().numpy().view(np.uint32)
            assert [envc.dec(r) for r in lw[::16]] == circuit.evaluate_plain(levels, bits).tolist()[::16], "layered circuit decrypts wrongly"
        per_gpu = -(-4096 // world)
        slots = 148 * 4
        configs["circuit_P0_g65536"] = {
            "workload": "config #4: 65536 gates uniform over {NAND, AND, XOR}, reference default parameters, sharded over all GPUs",
            "depth1": {"gates_per_s": G / (t1 * 1e-3), "ms": t1, "collectives": "NCCL scatter x2 + gather" if world > 1 else "none (1 GPU)"},
            "layered_16x4096": {"gates_per_s": 16 * 4096 / (t2 * 1e-3), "ms": t2,
                                "collectives": "one NCCL all-gather of the level's 4096 ciphertexts (11.8 MB) per level" if world > 1 else "none (1 GPU)",
                                "limiter": (f"wave quantisation: {per_gpu} gates per GPU per level on {slots} resident ciphertext slots "
                                            f"(148 CTAs x 4) = {-(-per_gpu // slots)} wave(s) filled to {100.0 * per_gpu / (slots * -(-per_gpu // slots)):.0f} %; "
                                            "16 levels are 16 dependent kernel chains of n = 722 steps; the all-gather is < 1 % of a level")}}
        envc.close()

    # ---- the one-process form of the same sharding (tfhe_mgpu C ABI, what a Rust host would call): rank 0 drives all GPUs
    mgpu_rec = None
    if world > 1 and not args.no_configs:
        barrier()
        if rank == 0:
            try:
                m = T.MultiGpuContext(p, world)
                mk = m.upload_key(E.bsk, E.ksk)
                tot = 65536
                cts, nu = E.batch(tot)
                hin = torch.from_numpy(cts.view(np.int32).copy()).pin_memory()
                hout = torch.empty((tot, row), dtype=i32).pin_memory()
                wall = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    m.bootstrap(mk, hin.numpy().view(np.uint32), E.tv, out=hout.numpy().view(np.uint32))
                    wall.append(1e3 * (time.perf_counter() - t0))
                r = hout.numpy().view(np.uint32)
                for i in range(0, tot, 1024):
                    assert E.dec(r[i]) == (i % nu) % E.pm
                din = hin.cuda()
                dwall = []
                for _ in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    dres = m.bootstrap(mk, din, E.tv)
                    dwall.append(1e3 * (time.perf_counter() - t0))
                assert np.array_equal(dres.cpu().numpy().view(np.uint32), r), "tfhe_mgpu device-pointer path differs from its host-pointer path"
                mgpu_rec = {"workload": f"P1: ONE batch of {tot} PBS through tfhe_mgpu_bootstrap_batch, one process driving {world} GPUs (the other ranks idle)",
                            "host_pointers": {"value": tot / (min(wall[1:]) * 1e-3), "unit": "PBS/s", "ms": min(wall[1:]), "timing": "host wall clock around the synchronous call, H2D/D2H per device inside"},
                            "device_pointers": {"value": tot / (min(dwall[1:]) * 1e-3), "unit": "PBS/s", "ms": min(dwall[1:]), "timing": "host wall clock; NCCL send/recv scatter + gather inside", "split_ms": m.last_timing()}}
                mk.free()
                m.close()
            except Exception as ex:   # reported, not fatal: the headline does not depend on this leg
                mgpu_rec = {"error": f"{type(ex).__name__}: {ex}"}
        dist.barrier(group=cpu_group)   # a host-side wait: an NCCL barrier would spin a kernel on the GPUs rank 0 is driving

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_pbs = world * B * args.steps
    value = total_pbs / (t_dev * 1e-3)
    e2e_value = total_pbs / (t_e2e * 1e-3)
    # roofline of the dominant kernel (blind rotation): algorithmic IMAD-class ops per launch / its device time
    ops_per_launch = B * p.n * c_cmux
    achieved = ops_per_launch / (t_br * 1e-3)
    P_, l_ = p.k + 1, p.pbs_levels
    M_ = p.N // 2
    # NTT path: 2 primes x 4 bytes per key word, one CTA per ciphertext, 3 CTAs per SM; FFT path: 2 limbs x 16 bytes per
    # pair of key words, the 3 (P1) or 4 (P0) ciphertexts of a CTA (one CTA per SM) share each key byte
    bsk_bytes = (p.n * P_ * l_ * 2 * P_ * M_ * 16) if fft_path else 2 * p.bsk_words * 4
    cts_per_sm = {9: 4, 10: 3}.get(p.glwe_poly_degree, 3) if fft_path else {9: 4, 10: 3, 11: 2}.get(p.glwe_poly_degree, 3)
    waves = -(-B // (148 * cts_per_sm))
    hbm_alg = bsk_bytes * waves + B * (p.n + 1) * 4 + B * p.glwe_words * 4
    bfly_per_launch = B * p.n * 2 * (p.N // 2) * p.glwe_poly_degree * (P_ * l_ + P_)
    traffic, traffic_source = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("preset") == args.preset and tj.get("batch") == B:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_source = f"profiles/{name}: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from a separate `ncu --set full` run of this command (not measured in this run)"
                break
    kernel_name = "pbs_fft_kernel (blind rotation, exact FP64-FFT path)" if fft_path else "pbs_kernel (blind rotation, 2-prime NTT path)"
    roofline = {"bound": "int32-imad", "achieved": achieved / 1e12, "peak": peaks["imad"] / 1e12, "unit": "T IMAD-class lane-ops/s",
                "frac": achieved / peaks["imad"], "traffic": traffic, "traffic_source": traffic_source, "kernel": kernel_name, "kernel_ms": t_br,
                "peak_source": "measured live: tfhe_measure_int_peak (dependent-free IMAD loop); imad_hi / imad_wide / Shoup-butterfly rates alongside",
                "peaks": {k: v / 1e12 for k, v in peaks.items()},
                "algorithmic_ops_per_launch": ops_per_launch,
                "butterflies_per_launch": bfly_per_launch,
                "frac_of_butterfly_ceiling": (bfly_per_launch / (t_br * 1e-3)) / peaks["shoup_butterfly"],
                "note": "IMAD.HI/IMAD.WIDE are half-rate, so a Shoup butterfly (1 HI + 2 IMAD) holds the FMA-heavy pipe for 8 cycles per warp, "
                        "not the 6 the 3-ops-per-butterfly accounting assumes; ncu FMA-heavy pipe utilisation is in profiles/",
                "hbm": {"algorithmic_bytes_per_launch": hbm_alg, "achieved_gbs": hbm_alg / (t_br * 1e-3) / 1e9, "peak_gbs": 6548.5,
                        "frac": hbm_alg / (t_br * 1e-3) / 1e9 / 6548.5, "note": "not the binding roofline (integer pipe binds by >100x)"}}
    if fft_path:
        # what the FFT kernel actually executes: FP64 operations (6 per forward butterfly, 8 per inverse butterfly, 4 per
        # complex multiply-accumulate, 1 per int->double conversion and per rounding), and the same count weighted by the
        # pipe cycles they occupy (a DFMA with three register operands holds the FP64 pipe 1.5x as long, measured)
        logm = p.glwe_poly_degree - 1
        bf = (M_ // 2) * logm
        rows = P_ * l_
        dp_ops = rows * bf * 6 + 2 * P_ * bf * 8 + rows * P_ * 2 * M_ * 4 + rows * p.N + 2 * P_ * p.N
        slow = rows * bf * 4 + 2 * P_ * bf * 2 + rows * P_ * 2 * M_ * 4     # three-register DFMAs among them
        r3 = fp64_peaks["dfma"] / fp64_peaks["dfma_3reg"]
        dp_cycles_equiv = dp_ops + slow * (r3 - 1.0)
        roofline.pop("butterflies_per_launch"); roofline.pop("frac_of_butterfly_ceiling")
        roofline["note"] = ("frac follows SURVEY 8(d): algorithmic IMAD-class work of the exact negacyclic product (numerator fixed, whatever "
                            "algorithm the kernel runs) over the measured IMAD peak.  This kernel computes the product on the FP64 pipe "
                            "(a separate pipe: the IMAD-bound NTT kernel reaches 0.41), see roofline.fp64")
        roofline["fp64"] = {"dp_ops_per_launch": B * p.n * dp_ops, "achieved_T": B * p.n * dp_ops / (t_br * 1e-3) / 1e12,
                            "peak_dfma_T": fp64_peaks["dfma"] / 1e12, "peak_dfma_3reg_T": fp64_peaks["dfma_3reg"] / 1e12,
                            "peak_fft_butterfly_T": fp64_peaks["fft_butterfly"] / 1e12,
                            "frac_of_dfma_peak": B * p.n * dp_ops / (t_br * 1e-3) / fp64_peaks["dfma"],
                            "frac_pipe_cycles": B * p.n * dp_cycles_equiv / (t_br * 1e-3) / fp64_peaks["dfma"],
                            "note": "frac_pipe_cycles weights three-register DFMAs by their measured issue cost; with the tensor-memory "
                                    "exchanges the kernels are bound by FP64 issue / dependent latency at 8-12 warps per SM rather than by "
                                    "the shared-memory pipe (ncu: profiles/r02_v15_pbs_fft_kernel_*_ncu_summary.txt), see roofline.smem"}
        # shared-memory / L1 data-pipe bytes the kernel moves per CMUX step and ciphertext (128-bit accesses, E = 8 points per
        # thread): transform exchanges, key rows read from the TMA ring, published/peer transformed rows, twiddles, digits
        # tensor-memory kernels (default): N = 512 has no shared-memory exchange and publishes rows in tensor memory; N = 1024 / 2048
        # keep one of the two exchanges per transform
        tmem = os.environ.get("TFHE_B200_FFT_TMEM", "1") != "0"
        xchg_ops = 32 if not tmem else {9: 0, 10: 16, 11: 16}.get(p.glwe_poly_degree, 32)     # LDS/STS.128 per transform and thread
        rows_in_smem = 0 if (tmem and p.glwe_poly_degree == 9) else 1
        E_, T_ = 8, M_ // 8
        per_thread = ((l_ + 2) * xchg_ops * 16 + l_ * P_ * 2 * E_ * 16 + rows_in_smem * (l_ * (P_ - 1) * E_ * 16 + l_ * E_ * 16) + (l_ + 2) * 2 * 16
                      + 2 * E_ * 4 * 2 + (l_ - 1) * 2 * E_ * 2 * 2 + 2 * E_ * 4 * 2)
        smem_bytes = B * p.n * per_thread * P_ * T_
        smem_peak = 148 * 128 * clocks["sm_mhz"] * 1e6 if clocks.get("sm_mhz") else 148 * 128 * 1.965e9
        roofline["smem"] = {"bytes_per_launch": smem_bytes, "achieved_TBs": smem_bytes / (t_br * 1e-3) / 1e12, "peak_TBs": smem_peak / 1e12,
                            "frac": smem_bytes / (t_br * 1e-3) / smem_peak,
                            "exchange": ("tensor memory (tcgen05.st/ld)" + ("" if p.glwe_poly_degree == 9 else " for the last stages, one shared-memory exchange per transform")) if tmem else "shared memory",
                            "note": "model of the LDS/STS bytes of one step; peak = 128 B/clk/SM x 148 SMs at the sampled SM clock; a barrier-synchronised "
                                    "pass loop (8 LDS.128 + 36 butterflies + 8 STS.128 per thread, 12 warps per SM) reaches 0.83 of it "
                                    "(profiles/r01_fp64_peak.json pass_loop_384thr_cycles)"}
    cfg = config_dict(args.preset, B, p)
    line = {"metric": "PBS/sec", "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": cfg,
            "data_path": ("inputs device-resident, one C-ABI call per step" if world == 1 else
                          f"the job's {total} ciphertexts live on rank 0: NCCL send/recv scatter -> {B} PBS per rank -> NCCL gather to rank 0, all inside the timed region"),
            "latency": {"ms_per_pbs_batch": t_dev / args.steps, "ms_batch1": latency_batch1_ms, "key_switch_ms": t_ks},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "PBS/s",
                    "h2d_bytes_per_step": int((host_root_in.numel() if world > 1 else host_in.numel()) * 4 + host_tv.numel() * 4 * (0 if world > 1 else 1)),
                    "d2h_bytes_per_step": int((host_root_out.numel() if world > 1 else host_out.numel()) * 4)},
            "roofline": roofline}
    if world > 1:
        line["replicas"] = {"value": total_pbs / (t_rep * 1e-3), "unit": "PBS/s", "note": "the same step without scatter/gather: every rank bootstraps a resident shard"}
    if configs:
        line["configs"] = configs
    if mgpu_rec:
        line["mgpu_c_abi"] = mgpu_rec
    if world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        o = orc.params(**fields)
        n_ct = 3 * cores
        cts = np.stack([E.enc(i % E.pm, i) for i in range(n_ct)])
        rate, dt, out = cpu_reference_rate(o, n_ct, cores, E.bsk, E.ksk, cts, E.tv)
        gpu_same = ctx.bootstrap(bk, cts, E.tv)
        assert np.array_equal(gpu_same, out), "GPU result differs from the CPU oracle on the baseline sample"
        line["cpu_baseline"] = {"value": rate, "unit": "PBS/s", "cores": cores, "kind": "port",
                                "sample": f"{n_ct} PBS of the same workload, one per host thread, {dt:.1f} s; C restatement of the reference's "
                                          "Rust path (Toeplitz O(N^2)); GPU output verified bit-identical on this sample"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _main_with_clean_stdout():
    """stdout carries exactly ONE JSON line: while the benchmark runs, file descriptor 1 points at stderr, so anything a
    library writes to stdout from C (NCCL prints its version banner there on some boxes) cannot precede the line."""
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            main()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    out = buf.getvalue()
    if out:
        sys.stdout.write(out)
        sys.stdout.flush()


if __name__ == "__main__":
    _main_with_clean_stdout()
