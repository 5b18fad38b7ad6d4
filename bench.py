#!/usr/bin/env python
"""bench.py — Mbp decoded per second on BASELINE.json config 2 (synthetic 50 kb human-composition windows,
--species=human ab initio), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--windows M] [--impl reference]

A "step" is one pass of the whole hot path (prep -> sweep -> backtrace -> pack) over the rank's batch of
windows.  `value` is timed with CUDA events on the library's launch stream with the inputs already in
HBM; `e2e` is the same metric through the public call (augb200_decode_batch) from host buffers, with the
host->device copy of the windows and the device->host copy of the paths inside the timed region.
Weak scaling: every rank decodes its own M windows (window index = rank*M + i), no data-path collective;
one NCCL gather of the final path arrays ends the e2e region.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/augustus, the unmodified
AUGUSTUS binary built by oracle/Makefile) on the box's host cores on a bounded sample of the same
windows.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW_LEN = 50000
DEFAULT_WINDOWS = 10000
STATES = 47
ALG_BYTES_PER_BASE = 0.25 + STATES * (8 + 8 + 4)          # 940 B/base (dense S=47 Viterbi)
METRIC = "Mbp decoded/sec"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------- reference arm
def ref_binary():
    exe = os.path.join(ROOT, "oracle", "_ref", "augustus")
    cfg = os.path.join(ROOT, "oracle", "_ref", "config")
    return (exe, cfg) if os.path.exists(exe) and os.path.isdir(cfg) else (None, None)


def run_reference_sample(n_windows, cores, start_index=0, extra_args=(), window_len=None, seqs=None, base_args=("--species=human", "--softmasking=0"),
                         one_per_process=False):
    """Decode n_windows synthetic windows (or the given sequences) with the unmodified reference, `cores` processes in parallel.
    Returns (Mbp/s, seconds)."""
    from augustus_b200 import synth
    exe, cfg = ref_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref/augustus is missing (run __graft_entry__.build() in the build container)")
    env = dict(os.environ, AUGUSTUS_CONFIG_PATH=cfg)
    wlen = window_len or WINDOW_LEN
    if seqs is not None:
        n_windows = len(seqs)
    with tempfile.TemporaryDirectory() as td:
        files = []
        nfile = n_windows if one_per_process else cores
        per = [[] for _ in range(nfile)]
        for i in range(n_windows):
            per[i % nfile].append(start_index + i)
        for c, idxs in enumerate(per):
            if not idxs:
                continue
            fa = os.path.join(td, "c%d.fa" % c)
            synth.write_fasta(fa, [seqs[i - start_index] if seqs is not None else synth.window(i, wlen) for i in idxs], ["w%d" % i for i in idxs])
            files.append(fa)
        t0 = time.perf_counter()
        procs = []
        if one_per_process and len(files) > cores:      # more windows than cores: a pool of `cores` running processes
            import concurrent.futures as cf
            def one(a):
                c, fa = a
                cmd = [exe] + list(base_args) + list(extra_args) + [fa]
                return subprocess.run(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode
            with cf.ThreadPoolExecutor(cores) as ex:
                if any(ex.map(one, enumerate(files))):
                    raise RuntimeError("reference process failed")
            dt = time.perf_counter() - t0
            return sum(len(s) for s in seqs) / 1e6 / dt if seqs is not None else n_windows * wlen / 1e6 / dt, dt
        for c, fa in enumerate(files):
            cmd = [exe] + list(base_args) + list(extra_args) + [fa]
            if os.path.exists("/usr/bin/taskset"):
                cmd = ["taskset", "-c", str(c % (os.cpu_count() or 1))] + cmd
            procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("reference process failed")
        dt = time.perf_counter() - t0
    if seqs is not None:
        return sum(len(x) for x in seqs) / 1e6 / dt, dt
    return n_windows * wlen / 1e6 / dt, dt


def chr2l_windows(window=200000, step=150000):
    """BASELINE.json configs[2]: examples/chr2L cut into 200 kb windows stepping 150 kb (157 windows).  The FASTA is
    reference DATA copied to oracle/_ref/data/ by oracle/Makefile (it travels with the built checker; never read from /root/reference)."""
    import gzip
    path = os.path.join(ROOT, "oracle", "_ref", "data", "chr2L.sm.fa.gz")
    if not os.path.exists(path):
        return None
    seq = "".join(l.strip() for l in gzip.open(path, "rt") if not l.startswith(">"))
    out = []
    a = 0
    while True:
        out.append(seq[a:a + window])
        if a + window >= len(seq):
            break
        a += step
    return out


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    use = cores
    per_step = use * 2
    exe, _ = ref_binary()
    base = {"metric": METRIC, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (LLDouble)", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "synthetic 50 kb human-composition windows, --species=human ab initio (BASELINE.json configs[1])",
                       "window_len": WINDOW_LEN, "windows_per_step": per_step}}
    if exe is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/augustus not built"}))
        return
    for w in range(min(args.warmup, 1)):
        run_reference_sample(use, use, 0)
    tot_t, tot_w = 0.0, 0
    for s in range(args.steps):
        _, dt = run_reference_sample(per_step, use, s * per_step)
        tot_t += dt
        tot_w += per_step
    v = tot_w * WINDOW_LEN / 1e6 / tot_t
    base.update({"value": v, "ms_per_step": 1e3 * tot_t / args.steps,
                 "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": use, "kind": "reference",
                                  "sample": "%d windows x 50 kb per step, one augustus process per core" % per_step},
                 "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# --------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=DEFAULT_WINDOWS, help="windows per GPU per step (default: BASELINE.json config 2)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-5 (posterior sampling) side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from augustus_b200 import Decoder, synth
    from tests import util

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from augustus_b200 import shard
    M = args.windows
    # global list of world*M windows, block-cyclic over ranks (window g -> rank g mod world): rank r decodes g = r, r+world, ...
    my_idx = shard.shard_indices(world * M, rank, world)
    wins = synth.windows_parallel_indices(my_idx, WINDOW_LEN)
    wins_b = [w.encode() for w in wins]
    bases = M * WINDOW_LEN
    # inputs of the config-4 side measurement (generated here, before CUDA work starts, by the same process pool)
    n4, l4 = 1000, 200000
    wins4_b = [w.encode() for w in synth.windows_parallel(n4, l4)] if (not args.no_secondary and rank == 0 and world == 1) else None
    dec = Decoder(util.blob_bytes(), local)
    stream = torch.cuda.ExternalStream(dec.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: inputs staged in HBM, K timed steps ----
    dec.stage(wins_b)
    for _ in range(args.warmup):
        dec.run_staged()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(args.steps):
        dec.run_staged()
        launches += dec.last_launch_count
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    paths = dec.fetch_staged()
    sweep_ms = dec.last_sweep_ms            # last run's sweep kernel, CUDA events around that launch
    assert all(p.status == 0 for p in paths)
    # ---- end-to-end through the public call, host buffers in, host paths out ----
    # untimed warm-up of the public call at full size (pinned + device buffers reach their final size; NCCL sets up its gather)
    n_st, status, logp, offset, pb, pe, pt, ptr = dec.decode_batch_raw(wins_b)
    if world > 1:
        shard.gather_to_rank0(shard.pack_paths(n_st, status, logp, offset, pb, pe, pt, ptr), device="cuda")
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    d2h = 0
    for _ in range(e2e_steps):
        n_st, status, logp, offset, pb, pe, pt, ptr = dec.decode_batch_raw(wins_b)
        assert not status.any()
        d2h = pb.nbytes + pe.nbytes + pt.nbytes + ptr.nbytes + 32 * len(n_st)
        if world > 1:   # the one gather of the final results (path arrays) over NCCL
            got = shard.gather_to_rank0(shard.pack_paths(n_st, status, logp, offset, pb, pe, pt, ptr), device="cuda")
            if rank == 0:
                assert sum(int(v[0]) for v in got) == world * M
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.stop_flag = True; sampler.join(timeout=2)

    t = torch.tensor([ms, e2e_s, sweep_ms], dtype=torch.float64, device="cuda")
    clk = sampler.summary()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    # every rank samples its own GPU: report the slowest median SM clock and the union of the throttle reasons, plus per-rank kernel times
    c = torch.tensor([-(clk["sm_mhz"] or 0)] + [1.0 if n in clk["reasons"] else 0.0 for n in names], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, dtype=torch.float64, device="cuda"); per_rank[rank] = sweep_ms
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        clk["sm_mhz"] = int(-c[0].item()); clk["reasons"] = [n for n, v in zip(names, c[1:].tolist()) if v > 0]
        clk["per_rank_sweep_ms"] = [round(x, 1) for x in per_rank.tolist()]
    ms, e2e_s, sweep_ms = (float(x) for x in t.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = ms / args.steps
    value = world * bases / 1e6 / (ms_per_step / 1e3)
    peak, how = peaks()
    achieved = ALG_BYTES_PER_BASE * bases / (sweep_ms / 1e3) / 1e9
    traffic = None
    try:   # DRAM bytes of the sweep kernel from the committed ncu --set full capture (per base, scaled to this launch)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_sweep_2368win_final.json")))
        traffic = prof["dram_bytes_per_base"] * bases
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 (Q23.40 fixed-point log scores)", "data": "synthetic",
        "config": {"workload": "%d synthetic 50 kb human-composition windows per GPU, --species=human ab initio, 47 states (BASELINE.json configs[1])" % M,
                   "window_len": WINDOW_LEN, "windows_per_gpu": M, "parallelism": "windows sharded over %d GPU(s), one warp per window" % world,
                   "l2": "inputs per step (%.0f MB DNA + %.0f GB of per-window workspace written and read by the kernels) exceed the 126 MB L2" % (bases / 1e6, bases * 332 / 1e9)},
        "e2e": {"value": world * bases / 1e6 / e2e_s, "unit": "Mbp/s", "h2d_bytes_per_step": bases, "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_sweep", "peak_source": how,
                     "note": "achieved = 940 B/base (dense S=47 figure, SURVEY.md 8d) x bases / sweep-kernel time (CUDA events around the sweep launches of the last step); traffic = ncu dram bytes/base of profiles/r1_sweep_2368win_final.json x bases; the sweep stores only non-zero cells and is bound by instruction delivery (GPC instruction cache at 94 % of its request rate, profiles/r1_sweep_fetch_bound.txt), not by HBM"},
        "clocks": clk,
        "sweep_ms": sweep_ms,
    }
    if not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            v, dt = run_reference_sample(cores * 2, cores)
            line["cpu_baseline"] = {"value": v, "unit": "Mbp/s", "cores": cores, "kind": "reference",
                                    "sample": "%d windows x 50 kb (%.1f s wall), one unmodified augustus process per core" % (cores * 2, dt)}
        except Exception as ex:   # the reference binary did not travel: time the oracle port instead
            orc = util.Oracle(util.blob_bytes())
            t0 = time.perf_counter(); k = 8
            for i in range(k):
                orc.viterbi(wins[i])
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": k * WINDOW_LEN / 1e6 / dt, "unit": "Mbp/s", "cores": 1, "kind": "port",
                                    "sample": "%d windows x 50 kb, oracle/ghmm_oracle.c, 1 thread (%s)" % (k, ex)}
    if not args.no_secondary and world == 1:
        # BASELINE.json configs[4]: --sample=100 --alternatives-from-sampling=true on 1000 x 50 kb windows (forward + sampling kernels)
        try:
            n5 = min(1000, M)
            dec.decode_batch_sampling_raw(wins_b[:8], 100)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            vit5, samp5 = dec.decode_batch_sampling_raw(wins_b[:n5], 100)
            torch.cuda.synchronize(); dt5 = time.perf_counter() - t0
            assert not samp5[1].any()
            sec = {"workload": "%d x 50 kb windows, --sample=100 --alternatives-from-sampling=true (Viterbi + forward + 99 sampled paths per window), 1 GPU, e2e from host buffers through augb200_decode_batch_sampling" % n5,
                   "value": n5 * WINDOW_LEN / 1e6 / dt5, "unit": "Mbp/s", "sampled_paths": int(len(samp5[0]))}
            if not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                v5, d5 = run_reference_sample(cores, cores, extra_args=("--sample=100", "--alternatives-from-sampling=true"))
                sec["cpu_baseline"] = {"value": v5, "unit": "Mbp/s", "cores": cores, "kind": "reference", "sample": "%d windows x 50 kb (%.1f s wall)" % (cores, d5)}
            line["secondary"] = {"config5_sampling": sec}
        except Exception as ex:
            line["secondary"] = {"config5_sampling": {"error": str(ex)}}
        # BASELINE.json configs[3]: --species=human --UTR=on over 1000 x 200 kb windows (71 states: UtrModel next to the coding states)
        try:
            dec.close()
            dec4 = Decoder(util.blob_bytes("human_utr"), local)
            dec4.decode_batch_raw(wins4_b[:16])
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out4 = dec4.decode_batch_raw(wins4_b)
            torch.cuda.synchronize(); dt4 = time.perf_counter() - t0
            assert not out4[1].any()
            sec4 = {"workload": "%d x 200 kb synthetic windows, --species=human --UTR=on --softmasking=0 (71 states), 1 GPU, e2e from host buffers through augb200_decode_batch" % n4,
                    "value": n4 * l4 / 1e6 / dt4, "unit": "Mbp/s", "sweep_ms": dec4.last_sweep_ms, "path_states": int(out4[0].sum())}
            if not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                v4, d4 = run_reference_sample(cores, cores, extra_args=("--UTR=on",), window_len=l4)
                sec4["cpu_baseline"] = {"value": v4, "unit": "Mbp/s", "cores": cores, "kind": "reference", "sample": "%d windows x 200 kb (%.1f s wall)" % (cores, d4)}
            line["secondary"]["config4_utr"] = sec4
            dec4.close()
        except Exception as ex:
            line["secondary"]["config4_utr"] = {"error": repr(ex)}
    if not args.no_secondary and world == 1:
        # BASELINE.json configs[2]: examples/chr2L --species=fly (defaults: UTR on, softmasking on, sample=100) in 200 kb windows
        try:
            w3 = chr2l_windows()
            if w3 is None:
                raise RuntimeError("oracle/_ref/data/chr2L.sm.fa.gz not present")
            w3b = [w.encode() for w in w3]
            dec3 = Decoder(util.blob_bytes("fly_softmask_utr"), local)
            dec3.decode_batch_sampling_raw(w3b[:4], 100)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            vit3, samp3 = dec3.decode_batch_sampling_raw(w3b, 100)
            torch.cuda.synchronize(); dt3 = time.perf_counter() - t0
            assert not vit3[1].any() and not samp3[1].any()
            mbp3 = sum(len(w) for w in w3) / 1e6
            sec3 = {"workload": "examples/chr2L (23.5 Mbp, soft-masked) in %d windows of 200 kb stepping 150 kb, --species=fly defaults (UTR on = 71 states, softmasking on, "
                                "sample=100: Viterbi + forward + 99 sampled paths per window), 1 GPU, e2e from host buffers through augb200_decode_batch_sampling" % len(w3),
                    "value": mbp3 / dt3, "unit": "Mbp/s", "windows": len(w3), "sweep_ms": dec3.last_sweep_ms, "path_states": int(vit3[0].sum()), "sampled_paths": int(len(samp3[0]))}
            if not args.no_cpu_baseline:
                cores = os.cpu_count() or 1
                # half the cores: one 200 kb fly window costs the reference ~46 s on a free core and the default run has to end within minutes
                sub = w3[: min(len(w3), max(1, cores // 2))]
                v3, d3 = run_reference_sample(0, cores, seqs=sub, base_args=("--species=fly",), one_per_process=True)
                sec3["cpu_baseline"] = {"value": v3, "unit": "Mbp/s", "cores": min(cores, len(sub)), "kind": "reference",
                                        "sample": "%d of the windows, one unmodified augustus --species=fly process per window (%.1f s wall)" % (len(sub), d3)}
            line.setdefault("secondary", {})["config3_chr2L"] = sec3
            dec3.close()
        except Exception as ex:
            line.setdefault("secondary", {})["config3_chr2L"] = {"error": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
