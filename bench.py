#!/usr/bin/env python
"""bench.py -- PBS/s of the batched programmable bootstrap (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--preset P1] [--batch 4096] [--impl reference]

A "step" is one pass of the hot path (blind rotation + sample extract + key switch) over one batch of
synthetic LWE ciphertexts.  Default workload = BASELINE.json configs[1]: batch 4096 PBS, N=1024,
n=630, identity test vector, one B200 (preset P1 of SURVEY.md 8(d)).  With N>1 (torchrun, one rank
per GPU) every rank bootstraps its own 4096-ciphertext shard with a replica of the keys (weak
scaling, no data-path collective); `value` = all ranks' PBS / max-over-ranks device time.

value  : inputs resident in HBM (CUDA torch tensors handed to the C-ABI as device pointers).
e2e    : the same call with pinned HOST buffers, H2D and D2H copies inside the timed region.
roofline: integer pipe (IMAD-class lane-ops, SURVEY 8(d) W_int) against the IMAD peak measured live
         on this GPU by tfhe_measure_int_peak; the HBM side is reported in roofline.hbm.
cpu_baseline / --impl reference: the C restatement of the reference's Rust path (oracle/, faithful
         Toeplitz O(N^2) algorithm, one PBS per host thread) -- the Rust crate cannot be built here.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


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


def cpu_reference_rate(p_fields, seconds_hint, threads, bsk, ksk, cts, tv):  # seconds_hint: number of PBS (0 = one per thread)
    """PBS/s of the oracle (faithful reference algorithm), one PBS per host thread."""
    from oracle import orc
    o = orc.params(**p_fields)
    orc.lib().orc_set_faithful_toeplitz(1)
    t0 = time.perf_counter()
    n_ct = seconds_hint or threads
    out = orc.bootstrap_batch(o, cts[:n_ct], bsk, ksk, tv, threads)
    dt = time.perf_counter() - t0
    orc.lib().orc_set_faithful_toeplitz(0)
    return n_ct / dt, dt, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--preset", default="P1")
    ap.add_argument("--batch", type=int, default=4096, help="ciphertexts per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import tfhe_research_b200 as T
    p = T.TfheParams.preset(args.preset)
    fields = {f: getattr(p, f) for f, _ in T.TfheParams._fields_}
    w_int, c_cmux, c_ks = w_int_per_pbs(p)
    workload = (f"{args.preset}: batch {args.batch} PBS per GPU, k={p.k} N={p.N} n={p.n} pbs(logB={p.pbs_log_base},l={p.pbs_levels}) "
                f"ks(logB={p.ks_log_base},l={p.ks_levels}) log_p={p.log_p}, identity test vector")
    cores = args.cpu_threads or (os.cpu_count() or 1)

    # synthetic inputs: real keys (seeded keygen) and real encryptions so results can be decrypted
    lwe_sk, glwe_sk, bsk, ksk = T.bootstrapping_key_gen(p, 0xB200)
    tv = T.construct_identity_test_vector(p)
    pm = 1 << p.log_p

    if args.impl == "reference":
        if rank != 0:
            return
        n_ct = cores
        cts = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(n_ct)])
        times = []
        for s in range(args.warmup + args.steps):
            rate, dt, out = cpu_reference_rate(fields, 0, cores, bsk, ksk, cts, tv)
            if s >= args.warmup:
                times.append(dt)
            if s == 0:
                assert all(T.decode_rounded(p, T.decrypt_lwe(lwe_sk, out[i])) == i % pm for i in range(n_ct))
        total = sum(times)
        value = n_ct * len(times) / total
        line = {"impl": "reference", "metric": "PBS/sec", "value": value, "unit": "PBS/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": workload, "note": "CPU arm: each step bootstraps `cores` ciphertexts of the same workload, one per host thread"},
                "cpu_baseline": {"value": value, "unit": "PBS/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_ct} PBS per step (one per host thread), C restatement of the reference's Rust path, Toeplitz O(N^2) algorithm"},
                "e2e": {"value": value, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PBS path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = T.Context(p, local_rank)
    stream = torch.cuda.Stream()        # a real (non-default) stream shared by torch events and the C-ABI ctx
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)  # so torch CUDA events bracket exactly the kernels of this ctx
    bk = ctx.upload_key(bsk, ksk)

    B = args.batch
    base = rank * B
    n_unique = min(B, 256)  # encrypt 256 distinct ciphertexts, tile to the batch (timing is data independent)
    uniq = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, (base + i) % pm), 1, base + i) for i in range(n_unique)])
    host_in = torch.from_numpy(np.tile(uniq, ((B + n_unique - 1) // n_unique, 1))[:B].view(np.int32).copy()).pin_memory()
    host_out = torch.empty((B, p.n + 1), dtype=torch.int32).pin_memory()
    host_tv = torch.from_numpy(tv.view(np.int32).copy()).pin_memory()
    d_in, d_tv = host_in.cuda(), host_tv.cuda()
    d_out = torch.empty((B, p.n + 1), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    peaks = ctx.measure_int_peak()
    fft_path = ctx.pbs_path == T.PATH_FFT
    fp64_peaks = ctx.measure_fp64_peak() if fft_path else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(kind, nsteps, timed):
        evs, kern_ms = [], []
        for _ in range(nsteps):
            flush.zero_()  # L2 flush between iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if kind == "device":
                ctx.bootstrap(bk, d_in, d_tv, out=d_out)
            else:
                ctx.bootstrap(bk, host_in, host_tv, out=host_out)
            e1.record(stream)
            evs.append((e0, e1))
            kern_ms.append(ctx.last_timing())
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], kern_ms

    # ---- device-resident (value)
    sampler = ClockSampler(local_rank)
    sampler.start()
    run("device", args.warmup, False)
    barrier()
    launches0 = ctx.launch_count
    sampler.mark()
    step_ms, kern = run("device", args.steps, True)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    # ---- end to end through the C-ABI with pinned host buffers
    run("host", 1, False)
    barrier()
    e2e_ms, _ = run("host", args.steps, True)
    barrier()

    # per-PBS latency: one ciphertext through the same call (n dependent CMUX steps + key switch)
    lat1 = []
    one_in, one_out = d_in[:1].contiguous(), torch.empty((1, p.n + 1), dtype=torch.int32, device="cuda")
    for _ in range(4):
        ctx.bootstrap(bk, one_in, d_tv, out=one_out)
        lat1.append(ctx.last_timing()["total_ms"])
    latency_batch1_ms = min(lat1[1:])

    # correctness of what was timed: decrypt a sample of the last outputs on this rank
    res = d_out.cpu().numpy().view(np.uint32)
    res_h = host_out.numpy().view(np.uint32)
    assert np.array_equal(res, res_h), "device-pointer and host-pointer paths disagree"
    for i in range(0, B, max(1, B // 64)):
        assert T.decode_rounded(p, T.decrypt_lwe(lwe_sk, res[i])) == (base + i % n_unique) % pm, f"PBS {i} decrypts wrongly"

    t_dev, t_e2e = sum(step_ms), sum(e2e_ms)
    t_br = sum(k["blind_rotate_ms"] for k in kern) / len(kern)
    t_ks = sum(k["key_switch_ms"] for k in kern) / len(kern)
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, t_br], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_br = tt.tolist()
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
    # practical ceiling of this instruction mix: butterflies only, at the measured register-resident
    # Shoup-butterfly rate (IMAD.HI and IMAD.WIDE issue at HALF the plain IMAD rate on B200)
    bfly_per_launch = B * p.n * 2 * (p.N // 2) * p.glwe_poly_degree * (P_ * l_ + P_)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("preset") == args.preset and tj.get("batch") == B:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    kernel_name = "pbs_fft_kernel (blind rotation, exact FP64-FFT path)" if fft_path else "pbs_kernel (blind rotation, 2-prime NTT path)"
    roofline = {"bound": "int32-imad", "achieved": achieved / 1e12, "peak": peaks["imad"] / 1e12, "unit": "T IMAD-class lane-ops/s",
                "frac": achieved / peaks["imad"], "traffic": traffic, "kernel": kernel_name, "kernel_ms": t_br,
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
                            "note": "frac_pipe_cycles weights three-register DFMAs by their measured issue cost; the binding resource "
                                    "of this kernel is the shared-memory data pipe, see roofline.smem"}
        # shared-memory / L1 data-pipe bytes the kernel moves per CMUX step and ciphertext (128-bit accesses, E = 8 points per
        # thread): transform exchanges, key rows read from the TMA ring, published/peer transformed rows, twiddles, digits
        E_, T_ = 8, M_ // 8
        per_thread = ((l_ + 2) * 32 * 16 + l_ * P_ * 2 * E_ * 16 + l_ * (P_ - 1) * E_ * 16 + l_ * E_ * 16 + (l_ + 2) * 2 * 16
                      + 2 * E_ * 4 * 2 + (l_ - 1) * 2 * E_ * 2 * 2 + 2 * E_ * 4 * 2)
        smem_bytes = B * p.n * per_thread * P_ * T_
        smem_peak = 148 * 128 * clocks["sm_mhz"] * 1e6 if clocks.get("sm_mhz") else 148 * 128 * 1.965e9
        roofline["smem"] = {"bytes_per_launch": smem_bytes, "achieved_TBs": smem_bytes / (t_br * 1e-3) / 1e12, "peak_TBs": smem_peak / 1e12,
                            "frac": smem_bytes / (t_br * 1e-3) / smem_peak,
                            "note": "peak = 128 B/clk/SM x 148 SMs at the sampled SM clock; a barrier-synchronised pass loop of the same shape "
                                    "(8 LDS.128 + 36 butterflies + 8 STS.128 per thread, 12 warps per SM) reaches 0.83 of it "
                                    "(profiles/r01_fp64_peak.json pass_loop_384thr_cycles)"}
    line = {"metric": "PBS/sec", "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": workload, "l2": "256 MB flush write between timed iterations", "keys": "replicated per GPU",
                       "latency_ms_per_pbs_batch": t_dev / args.steps, "latency_ms_batch1": latency_batch1_ms, "key_switch_ms": t_ks},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "PBS/s", "h2d_bytes_per_step": int(host_in.numel() * 4 + host_tv.numel() * 4),
                    "d2h_bytes_per_step": int(host_out.numel() * 4)},
            "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        n_ct = 3 * cores
        cts = np.stack([T.encrypt_lwe_plaintext(p, lwe_sk, T.encode_message(p, i % pm), 1, i) for i in range(n_ct)])
        rate, dt, out = cpu_reference_rate(fields, n_ct, cores, bsk, ksk, cts, tv)
        gpu_same = ctx.bootstrap(bk, cts, tv)
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
