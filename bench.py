#!/usr/bin/env python
"""bench.py -- mapped Gbp/s of the seed-chain-align hot path (BASELINE.json metric) on synthetic long reads.

One "step" = one pass of the whole hot path (K1 sketch .. K8 base alignment + ds) over one batch of reads.
Workload at N=1 (BASELINE.json configs[1]): test/MT.gfa <- 10 000 x 10 kb ONT-error reads, -cx lr -c.
With N>1 every rank maps its own 10 000-read batch (seed + rank) against a replicated index: weak scaling, no
data-path collective; the only collective is the all-gather of per-rank GAF byte counts that fixes output offsets.

  value   = bases / device time of the five stage kernels (CUDA events inside libmgb200, inputs resident in HBM)
  e2e     = bases / wall time of mg_map_batch() from HOST buffers (H2D, kernels, D2H, result assembly) + GAF formatting
  roofline= chaining kernel (K4/K5, `k_stage<1>`): algorithmic bytes (16 B/seed in + 16 B/anchor out + 8 B/chain) / its
            event time, against MEASURED_PEAKS.json hbm_gbs
  --impl reference : the unmodified reference CLI (oracle/_ref/minigraph -t <all cores>) on a bounded sample
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

FIX = os.path.join(REPO, "tests", "golden", "fixtures")
MGSIM = os.path.join(REPO, "tools", "mgsim")
REF_BIN = os.path.join(REPO, "oracle", "_ref", "minigraph")
MT_WALKS = [">MTh0>MTh4001>MTh4502>MTh9505>MTh13014>MTh13516", ">MTh0<MTo3426>MTh4502>MTo8961>MTh9505>MTh13516"]

N_READS, READ_LEN = 10000, 10000
CPU_SAMPLE_READS = 3000


def make_workload_c3(tmp, rank, n_reads, graph_len=5000000, n_hap=8, read_len=15000):
    """synthetic MHC-scale rGFA (5 Mb backbone, 8 haplotypes with SVs, seed 7) and ONT-error reads (seed 5+rank)."""
    prefix = os.path.join(tmp, "mhc")
    if not os.path.exists(prefix + ".gfa"):
        subprocess.run([MGSIM, "graph", "-l", str(graph_len), "-n", str(n_hap), "-s", "7", "-o", prefix], check=True, stderr=subprocess.DEVNULL)
    reads = os.path.join(tmp, "mhc.reads.%d.fa" % rank)
    subprocess.run([MGSIM, "reads", "-i", prefix + ".hap.fa", "-n", str(n_reads), "-l", str(read_len), "-e", "ont", "-s", str(5 + rank), "-o", reads],
                   check=True, stderr=subprocess.DEVNULL)
    return prefix + ".gfa", reads


def make_workload(tmp, rank, n_reads=N_READS):
    hap = os.path.join(tmp, "mt.hap.fa")
    reads = os.path.join(tmp, "mt.reads.%d.fa" % rank)
    cmd = [MGSIM, "walk", "-g", os.path.join(FIX, "MT.gfa"), "-o", hap]
    for w in MT_WALKS:
        cmd += ["-w", w]
    subprocess.run(cmd, check=True)
    subprocess.run([MGSIM, "reads", "-i", hap, "-n", str(n_reads), "-l", str(READ_LEN), "-e", "ont", "-s", str(11 + rank), "-o", reads],
                   check=True, stderr=subprocess.DEVNULL)
    return os.path.join(FIX, "MT.gfa"), reads


def read_fasta(fn):
    names, seqs = [], []
    with open(fn, "rb") as f:
        data = f.read().split(b"\n")
    for i in range(0, len(data) - 1, 2):
        names.append(data[i][1:])
        seqs.append(data[i + 1])
    return names, seqs


def run_reference_cli(gfa, fasta, threads):
    """Mapping-phase seconds of the reference (last 'mapped' stamp minus the 'indexed' stamp)."""
    t0 = time.perf_counter()
    p = subprocess.run([REF_BIN, "-cx", "lr", "-t", str(threads), gfa, fasta], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True)
    wall = time.perf_counter() - t0
    log = p.stderr.decode()
    t_idx = re.findall(r"\[M::mg_index::([0-9.]+)\*", log)
    t_map = re.findall(r"\[M::worker_pipeline::([0-9.]+)\*[0-9.]+\] mapped", log)
    if t_idx and t_map:
        return float(t_map[-1]) - float(t_idx[-1]), wall
    return wall, wall


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.rows)}


def chain_traffic():
    """DRAM bytes of one k_chain launch from the committed `ncu --set full` digest (profiles/r01_ncu_k_chain.txt), or None."""
    try:
        rd = wr = None
        with open(os.path.join(REPO, "profiles", "r01_ncu_k_chain.txt")) as f:
            for ln in f:
                t = ln.rstrip("\n").split("\t")
                if len(t) == 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v = float(t[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[t[2]]
                    if t[0].startswith("dram__bytes_read"):
                        rd = v
                    else:
                        wr = v
        return rd + wr if rd is not None and wr is not None else None
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mgb200")
    ap.add_argument("--reads", type=int, default=N_READS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"], help="c2: test/MT.gfa <- 10 kb reads (the metric's configuration, default); c3: synthetic MHC-scale rGFA <- 15 kb reads")
    ap.add_argument("--check", type=int, default=1000, help="after the timed region, compare the GAF text of the first N reads with the reference binary (oracle/_ref/minigraph), byte for byte; 0 to skip")
    a = ap.parse_args()
    # stdout carries exactly one JSON line: everything libraries print to fd 1 (NCCL's version banner, ...) goes to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    tmp = tempfile.mkdtemp(prefix="mgb_bench_")
    ncores = os.cpu_count() or 1
    workload = "test/MT.gfa <- %d x %d bp synthetic ONT-error reads (4%% sub, 3%% del, 3%% ins; mgsim seed 11+rank), -cx lr -c" % (a.reads, READ_LEN)
    if a.workload == "c3":
        workload = "synthetic MHC-scale rGFA (5 Mb backbone, 8 haplotypes with SVs, mgsim seed 7) <- %d x 15000 bp ONT-error reads (seed 5+rank), -cx lr -c" % a.reads

    if a.impl == "reference":
        if rank != 0:
            return
        gfa, fa = make_workload_c3(tmp, 0, CPU_SAMPLE_READS) if a.workload == "c3" else make_workload(tmp, 0, CPU_SAMPLE_READS)
        names, seqs = read_fasta(fa)
        bases = sum(len(s) for s in seqs)
        for _ in range(min(a.warmup, 1)):
            run_reference_cli(gfa, fa, ncores)
        ts = [run_reference_cli(gfa, fa, ncores)[0] for _ in range(a.steps)]
        t = sum(ts) / len(ts)
        v = bases / t / 1e9
        sample = "%d of the %d reads (%.1f Mbp), mapping phase of `minigraph -cx lr -t %d`" % (CPU_SAMPLE_READS, a.reads, bases / 1e6, ncores)
        json_out.write(json.dumps({
            "impl": "reference", "metric": "mapped Gbp/s (-cx lr)", "value": v, "unit": "Gbp/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/u8 (fp32 chain penalties)",
            "data": "synthetic", "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "Gbp/s", "cores": ncores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }) + "\n")
        json_out.flush()
        return

    import torch
    import torch.distributed as dist
    from minigraph_b200 import capi, options
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.load_product()  # applies MGB_PARAMS (engine switches for experiments)
    lib.mgb_set_param(b"device", local_rank)
    # host threads for packing, result assembly and GAF text: this rank's share of the box (two pools run at once)
    host_threads = max(4, min(48, ncores // (2 * world)))
    lib.mgb_set_param(b"host_threads", host_threads)
    gfa, fa = make_workload_c3(tmp, rank, a.reads) if a.workload == "c3" else make_workload(tmp, rank, a.reads)
    names, seqs = read_fasta(fa)
    n = len(seqs)
    bases = sum(len(s) for s in seqs)
    g = lib.mgb_gfa_read(gfa.encode())
    io, mo = options.opt_set("lr", cigar=True)
    gi = lib.mg_index(g, C.byref(io), 1, C.byref(mo))
    assert gi, lib.mgb_last_error()
    qlens = (C.c_int * n)(*[len(s) for s in seqs])
    cseqs = (C.c_char_p * n)(*seqs)
    cnames = (C.c_char_p * n)(*names)
    # two result sets: the GAF text of batch i is written by a second host thread while batch i+1 is being mapped, the
    # way the reference's kt_pipeline overlaps its output step with the mapping step of the next mini-batch (gmap.c:176-177)
    gcs2 = [(C.POINTER(capi.mg_gchains_t) * n)(), (C.POINTER(capi.mg_gchains_t) * n)()]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    st = capi.mgb_stats_t()
    gaf_bytes = [0]
    host = [0.0] * 6
    import queue
    out_q, done_q = queue.Queue(), queue.Queue()
    gaf_buf, gaf_cap = C.c_void_p(0), C.c_size_t(0)  # the output buffer is handed back to the writer every batch

    def output_worker():
        while True:
            k = out_q.get()
            if k is None:
                return
            t0 = time.perf_counter()
            ln = C.c_size_t(0)
            lib.mgb_write_gaf_batch(g, n, gcs2[k], qlens, cnames, mo.flag, host_threads, C.byref(gaf_buf), C.byref(ln), C.byref(gaf_cap))
            gaf_bytes[0] = ln.value
            lib.mgb_free_batch(n, gcs2[k])
            host[5] += (time.perf_counter() - t0) * 1e3
            done_q.put(k)

    worker = threading.Thread(target=output_worker, daemon=True)
    worker.start()
    free_sets = [0, 1]
    step_no = [0]

    def step():
        flush.fill_(1)  # evict L2 (126 MB) between steps
        torch.cuda.synchronize()
        while not free_sets:
            free_sets.append(done_q.get())
        k = free_sets.pop(0)
        rc = lib.mg_map_batch(gi, n, qlens, cseqs, cnames, gcs2[k], C.byref(mo))
        assert rc == 0, lib.mgb_last_error()
        lib.mgb_get_stats(gi, C.byref(st))
        host[0] += st.t_pack_ms; host[1] += st.t_h2d_ms; host[2] += st.t_d2h_ms; host[3] += st.t_asm_ms; host[4] += st.t_host_ms
        out_q.put(k)
        step_no[0] += 1
        return (st.t_seed_ms, st.t_chain_ms, st.t_align_ms, st.t_wfa_ms, st.t_finish_ms), st.t_dev_span_ms, int(st.n_slots)

    def drain():
        while len(free_sets) < 2:
            free_sets.append(done_q.get())

    for _ in range(a.warmup):
        step()
    drain()
    host[:] = [0.0] * 6
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kern, stage = [], [0.0] * 5
    launches = 0
    t_begin = time.perf_counter()
    for _ in range(a.steps):
        ks, span, n_slots = step()
        kern.append(span)
        for i in range(5):
            stage[i] += ks[i]
        launches += st.n_launches
    drain()  # the last batch's GAF text is part of the job
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_begin
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # one extra, untimed-for-the-metric step with a single sub-batch: kernels run back to back on one stream, so the CUDA-event
    # times of the stages are per-kernel durations (in the timed steps the sub-batches overlap and share the SMs)
    overlapped_stage = [x / a.steps for x in stage]
    host_timed = list(host)  # the host-side sums of the timed steps
    lib.mgb_set_param(b"slots", 1)
    serial_stage, serial_span, _ = step()
    drain()
    stage = [x * a.steps for x in serial_stage]
    kernel_ms = {k: round(st.t_kernel_ms[i], 3) for i, k in enumerate(capi.KERNEL_NAMES) if k != "k_index_sketch"}
    t_kern = sum(kern) / 1e3
    # the one collective of the path: per-rank GAF byte counts -> output offsets
    from minigraph_b200 import dist as mdist
    my_off, counts = mdist.gaf_offsets(gaf_bytes[0], device="cuda")
    offsets = [sum(counts[:i]) for i in range(len(counts))]
    if world > 1:
        tm = torch.tensor([t_kern, t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        t_kern, t_wall = float(tm[0].item()), float(tm[1].item())
    total_bases = bases * world
    value = total_bases * a.steps / t_kern / 1e9
    e2e = total_bases * a.steps / t_wall / 1e9
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    check = None
    if a.check > 0 and os.path.exists(REF_BIN):  # untimed: the first reads once more, GAF text against the reference binary
        try:
            m = min(a.check, n)
            cfa = os.path.join(tmp, "check.fa")
            with open(cfa, "wb") as f:
                for nm, sq in zip(names[:m], seqs[:m]):
                    f.write(b">" + nm + b"\n" + sq + b"\n")
            want = subprocess.run([REF_BIN, "-cx", "lr", "-t", str(ncores), gfa, cfa], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            gcs_c = (C.POINTER(capi.mg_gchains_t) * m)()
            rc = lib.mg_map_batch(gi, m, qlens, cseqs, cnames, gcs_c, C.byref(mo))
            assert rc == 0, lib.mgb_last_error()
            buf, ln = C.c_void_p(0), C.c_size_t(0)
            lib.mgb_write_gaf_batch(g, m, gcs_c, qlens, cnames, mo.flag, host_threads, C.byref(buf), C.byref(ln), None)
            got = C.string_at(buf, ln.value)
            C.CDLL(None).free(buf)
            lib.mgb_free_batch(m, gcs_c)
            check = {"reads": m, "gaf_bytes": len(want), "identical": got == want}
        except Exception as e:  # the measured line is still worth printing; the failure is reported in it
            check = {"reads": 0, "identical": False, "error": repr(e)[:200]}

    peak, peak_src = hbm_peak()
    chain_bytes = 16.0 * st.n_seeds + 16.0 * st.n_anchors_out + 8.0 * st.n_chains_out
    t_chain_avg = stage[1] / a.steps / 1e3
    achieved = chain_bytes / t_chain_avg / 1e9 if t_chain_avg > 0 else 0.0
    # CPU baseline: the unmodified reference on a bounded sample of the same reads, all host cores
    cpu = None
    if os.path.exists(REF_BIN) and not a.no_cpu:
        sfa = os.path.join(tmp, "cpu_sample.fa")
        with open(sfa, "wb") as f:
            for nm, s in zip(names[:CPU_SAMPLE_READS], seqs[:CPU_SAMPLE_READS]):
                f.write(b">" + nm + b"\n" + s + b"\n")
        sb = sum(len(s) for s in seqs[:CPU_SAMPLE_READS])
        run_reference_cli(gfa, sfa, ncores)
        tcpu = min(run_reference_cli(gfa, sfa, ncores)[0] for _ in range(2))
        cpu = {"value": sb / tcpu / 1e9, "unit": "Gbp/s", "cores": ncores, "kind": "reference",
               "sample": "first %d reads of the batch (%.1f Mbp), mapping phase of `oracle/_ref/minigraph -cx lr -t %d`" % (min(CPU_SAMPLE_READS, n), sb / 1e6, ncores)}
    out = {
        "metric": "mapped Gbp/s (-cx lr)", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t_kern / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64/u8 (fp32 chain penalties, bit-exact)", "data": "synthetic",
        "config": {"workload": workload, "reads_per_gpu": n, "bases_per_gpu": bases, "l2": "512 MiB flush buffer written between steps", "engine_params": capi.env_params(),
                   "parallelism": "reads sharded one batch per GPU, index replicated; all-gather of GAF byte counts only",
                   "gaf_offsets": offsets},
        "sub_batches": n_slots, "host_threads": host_threads, "host_cores": ncores,
        "stage_ms_note": "stage_ms_per_step comes from one extra step run as a single sub-batch (kernels serialised on one stream, %.1f ms device span); in the timed steps %d sub-batches overlap and the summed per-stream stage times were %s" % (serial_span, n_slots, ["%.1f" % x for x in overlapped_stage]),
        "kernel_ms": kernel_ms,
        "stage_ms_per_step": {"seed(K1-K3)": stage[0] / a.steps, "chain(K4-K5)": stage[1] / a.steps, "gchain+plan(K6-K7)": stage[2] / a.steps,
                              "wfa_jobs(K8a)": stage[3] / a.steps, "finish(K8b cigar+ds)": stage[4] / a.steps, "wfa_jobs_per_step": int(st.n_jobs), "wfa_jobs_tier2": int(st.n_jobs_mid), "wfa_jobs_tier3": int(st.n_jobs_big)},
        "e2e": {"value": e2e, "unit": "Gbp/s", "ms_per_step": t_wall / a.steps * 1e3,
                "h2d_bytes_per_step": int(bases + 16 * n + 16 * n), "d2h_bytes_per_step": int(st.out_bytes + 48 * n + 96 * n),
                "includes": "wall clock of K x (pack + H2D of reads, all kernels, D2H of result blobs, mg_gchains_t assembly) with the GAF text (%d bytes/step) of batch i written by a second host thread during batch i+1 (the reference's kt_pipeline does the same, gmap.c:176), plus the last batch's GAF; L2 flush between steps included" % gaf_bytes[0]},
        "host_ms_per_step": {"pack": host_timed[0] / a.steps, "h2d": host_timed[1] / a.steps, "d2h": host_timed[2] / a.steps, "assemble": host_timed[3] / a.steps,
                             "mg_map_batch_total": host_timed[4] / a.steps, "gaf_text(second thread)": host_timed[5] / a.steps},
        "device_cycles_last_step": {k: (int(st.prof[i]) >> 16 if k in ("wfa_max_cyc", "gwfa_max_cyc") else int(st.prof[i])) for i, k in enumerate(capi.PROF_NAMES)},
        "slowest_units": {"gap_len": int(st.prof[capi.PROF_NAMES.index("wfa_max_cyc")]) & 0xffff, "bridge_len": int(st.prof[capi.PROF_NAMES.index("gwfa_max_cyc")]) & 0xffff},
        "wfa_tier_routing": {"skip_tier1_at": int(st.skip1_len), "skip_tier2_at": int(st.skip2_len)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_chain (linear chaining: mg_lchain_dp/rmq + backtrack + compaction)", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": chain_traffic(), "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_chain launch on this workload, profiles/r01_ncu_k_chain.txt",
                     "algorithmic_bytes_per_launch": chain_bytes, "launch_ms": t_chain_avg * 1e3, "peak_source": peak_src,
                     "note": "the chaining kernel is bound by dependent-access latency and instruction issue, not by HBM bandwidth (DESIGN.md section 4)",
                     "bytes_model": "16 B x %d seeds in + 16 B x %d anchors out + 8 B x %d chains" % (st.n_seeds, st.n_anchors_out, st.n_chains_out)},
        "cpu_baseline": cpu, "parity_check": check,
        "clocks": sampler.summary(),
    }
    json_out.write(json.dumps(out) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
