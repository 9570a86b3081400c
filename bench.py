#!/usr/bin/env python
"""Benchmark of qboot's data-parallel hot path on B200 -- BASELINE.json metric: "gBlock tables/s (lambda = 19,
n_Max = 400, 1024-bit) + full PMP build time vs CPU".

  python bench.py --gpus N --steps K --warmup W            this engine (CUDA, through the C ABI / the C++ facade)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation (oracle/_ref)

M1 -- block tables per second.  A "step" is one pass of the table pipeline over the scan workload of BASELINE config 5:
the mixed-correlator Ising key set (spins 0..50, numax = 20, three (S, P) classes) repeated over a grid of
(Delta_sigma, Delta_epsilon); --total-tables keys per step IN TOTAL, dealt round-robin by cost class (spin 0 / spin > 0)
to the ranks: STRONG scaling.  With N > 1 every step ends with the gather of all tables into rank 0's HBM over
NVLink (NCCL grouped send/recv through torch.distributed), inside the timed region.
  value     tables/s, inputs resident in HBM (qb_gblock_plan done; qb_gblock_run [+ gather] timed with CUDA events on the
            launching stream; barrier + synchronize on both sides; max over ranks)
  e2e       tables/s through qb_gblock_batch with HOST buffers in and out (planning, H2D, kernels, D2H all timed)
  roofline  whole step against the MEASURED limb-product rate of this GPU (qb_imad_peak, IMAD.WIDE.U32): the path is
            multiply-bound (>= 8 MAC per HBM byte against a machine balance of ~1), so the bound is "imad", not
            hbm/tensor; per-stage fractions in `stages`; `traffic` = ncu dram bytes of the dominant stage per launch
  cpu_baseline  oracle/_ref (the unmodified reference, MPFR) on this box's host cores, bounded strided sample, rank 0, N = 1
M2 -- full PMP build (convert -> create_input -> write), N = 1 only: `pmp_build` holds wall seconds per phase of the
  single-correlator (1 289 tables) and mixed-correlator (6 254 tables, 407 MB of SDPB text) north-star builds through the
  C++ facade (qboot::Context ... SDPBInput::write on the GPU engine), beside the reference's own build of the same
  model on all host cores, and whether the two directories are byte-identical.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PREC, N_MAX, LAM, DIM, NUMAX, MAXSPIN = 1024, 400, 19, "3", 20, 50
METRIC = "gblock_tables_per_s"
UNIT = "tables/s"
WORKLOAD = ("gBlock tables, Context(n_Max=400, lambda=19, dim=3), 1024-bit, mixed-Ising key set (spins 0-50, numax=20) swept over a "
            "(Delta_sigma, Delta_epsilon) grid; N = 1 also: full PMP build of the single- and mixed-correlator Ising problems")
TOTAL_TABLES = 392768  # 8 x 49 096: 64 grid points of the 13 x 13 scan (a fixed total for every N: strong scaling)


_RESULT_FD = None


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_shard(rank: int, world: int, total: int):
    """this rank's share of the scan workload (deterministic): the keys of `total` tables, sorted by cost class
    (spin 0 tables run a shorter recurrence than spin > 0) and dealt round-robin -- SURVEY section 8(e)"""
    from qboot_b200 import workload as wl

    keys = wl.sweep_keys(PREC, LAM, NUMAX, MAXSPIN, total)
    if world > 1:
        order = sorted(range(len(keys)), key=lambda i: (keys[i][0] != 0, i))
        keys = [keys[i] for i in order[rank::world]]
    return wl.to_records(keys, PREC)


def total_work(spin, stage="total"):
    from qboot_b200 import workload as wl

    w0 = wl.work_per_table(N_MAX, LAM, 32, 0)[stage]
    w1 = wl.work_per_table(N_MAX, LAM, 32, 1)[stage]
    n0 = int((spin == 0).sum())
    return n0 * w0 + (len(spin) - n0) * w1


def strided_sample(n: int, count: int):
    return np.arange(0, n, max(1, n // count))[:count]


def run_reference(args, rank, world):
    """the reference's own CPU implementation of the path (oracle/_ref), all host threads, bounded sample per step"""
    if rank != 0:
        return 0
    from oracle import ref

    if not ref.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libqboot_ref.so not built (needs /root/reference at build time)"})
        return 0
    cores = host_cores()
    sample = int(args.ref_tables) if args.ref_tables else max(64, 16 * cores)
    spin, delta, S, P = make_shard(0, 1, 49096)
    idx = strided_sample(len(spin), sample)  # every sector, spin and (S, P) class of the key set
    spin, delta, S, P = spin[idx], delta[idx], S[idx], P[idx]
    for _ in range(args.warmup):
        ref.time_gblock(PREC, N_MAX, LAM, DIM, delta[:cores], spin[:cores], S[:cores], P[:cores], cores)
    t = 0.0
    for _ in range(args.steps):
        t += ref.time_gblock(PREC, N_MAX, LAM, DIM, delta, spin, S, P, cores)
    value = len(idx) * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "mpfr-1024 (f1024)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "tables_per_step": int(len(idx))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{len(idx)} tables per step, strided over the whole key set, through qboot::gBlock on {cores} host threads (reference task fan-out)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_pmp:
        line["pmp_build"] = {"single": ref_pmp(0, cores), "mixed": ref_pmp(1, cores)}
    emit(line)
    return 0


def ref_pmp(model: int, cores: int, outdir=None):
    from oracle import ref

    d = outdir or tempfile.mkdtemp(prefix="qb_ref_pmp_")
    t = time.perf_counter()
    r = ref.build_pmp(PREC, N_MAX, LAM, DIM, model, 1 if model else 0, True, numax=NUMAX, maxspin=MAXSPIN, threads=cores, outdir=d)
    r["wall_s"] = time.perf_counter() - t
    r["threads"] = cores
    if outdir is None:
        shutil.rmtree(d, ignore_errors=True)
    return r


def pmp_leg(cores: int, with_reference: bool):
    """M2: convert -> create_input -> write through the C++ facade, cold and warm, beside the reference on this box"""
    import filecmp

    import qboot_b200 as qb

    out = {}
    for name, model in (("single", 0), ("mixed", 1)):
        root = tempfile.mkdtemp(prefix="qb_pmp_")
        runs = []
        for rep in range(4):  # the first call pays module loading and page-locking; the best of the other three is reported
            d = os.path.join(root, f"ours{rep}")
            t = time.perf_counter()
            r = qb.build_pmp(PREC, N_MAX, LAM, DIM, model, 1 if model else 0, True, numax=NUMAX, maxspin=MAXSPIN, threads=cores, outdir=d)
            r["wall_s"] = time.perf_counter() - t
            runs.append(r)
        best = min(runs[1:], key=lambda r: r["wall_s"])
        rec = {"wall_s": best["wall_s"], "convert_s": best["convert"], "create_input_s": best["create_input"], "write_s": best["write"],
               "first_call_wall_s": runs[0]["wall_s"], "constraints": best["constraints"], "N": best["N"],
               "sdpb_bytes": sum(os.path.getsize(os.path.join(root, "ours1", f)) for f in os.listdir(os.path.join(root, "ours1"))),
               "host_threads": cores}
        if with_reference:
            rd = os.path.join(root, "ref")
            rr = ref_pmp(model, cores, rd)
            files = sorted(os.listdir(rd))
            differing = [f for f in files if not filecmp.cmp(os.path.join(rd, f), os.path.join(root, "ours1", f), shallow=False)]
            rec["reference"] = {"wall_s": rr["wall_s"], "convert_s": rr["convert"], "create_input_s": rr["create_input"], "write_s": rr["write"],
                                "tables": rr["tables"], "threads": cores, "kind": "unmodified reference (oracle/_ref), all host cores"}
            rec["tables"] = rr["tables"]
            rec["files"] = len(files)
            rec["byte_identical_to_reference"] = (differing == [] and sorted(os.listdir(os.path.join(root, "ours1"))) == files)
            rec["speedup_vs_reference"] = rr["wall_s"] / best["wall_s"]
        shutil.rmtree(root, ignore_errors=True)
        out[name] = rec
    return out


def ncu_traffic(n_tables):
    """DRAM bytes of one step from the committed ncu --set full capture (profiles/r2_traffic.json: dram__bytes_read.sum +
    dram__bytes_write.sum of every pipeline kernel, per table), scaled to this run's tables per step; None when no capture
    is recorded."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        return rec["step_dram_bytes_per_table"] * n_tables
    except Exception:
        return None


class _DevBytes:
    """a raw device allocation of the engine as something torch can wrap (no copy)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--total-tables", type=int, default=int(os.environ.get("QB_BENCH_TABLES", str(TOTAL_TABLES))))
    ap.add_argument("--ref-tables", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pmp", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything else a library may print there (NCCL's version banner,
    # torchrun notices) is sent to stderr for the lifetime of the process
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rules: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import qboot_b200 as qb

    eng = qb.Engine(PREC, device=local).context(N_MAX, LAM, DIM)
    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    total = args.total_tables - args.total_tables % world
    spin, delta, S, P = make_shard(rank, world, total)
    n = len(spin)
    W = eng.W
    table_bytes = eng.dim_M * W * 4
    # pinned host staging for the end-to-end leg
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_spin, h_delta, h_S, h_P = pin(spin.astype(np.uint32).view(np.int32)), pin(delta.view(np.int32)), pin(S.view(np.int32)), pin(P.view(np.int32))
    h_out = torch.empty((n, eng.dim_M, W), dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs already in HBM ---------------------------------------------------------------------
    eng.plan(spin, delta, S, P)
    peak = eng.imad_peak(4096)
    eng.set_timing(False)
    # ---- the gather (N > 1): every rank's tables end up in ONE buffer in rank 0's HBM.  Rank 0 allocates it and exports a
    # CUDA IPC handle; every rank opens it and names its slice as the sink of its runs: each finished chunk is then copied
    # there by the copy engine over NVLink underneath the kernels of the later chunks (no collective on the data path;
    # torch.distributed only carries the 64-byte handle, the barriers and the timing reduction).  QB_GATHER=nccl selects the
    # plain alternative, torch.distributed.gather after the last kernel.
    gather_mode = os.environ.get("QB_GATHER", "peer") if world > 1 else "none"
    shard_bytes = n * table_bytes
    gathered, mine, sink_base, sink_local = None, None, None, None
    if world > 1 and gather_mode == "peer":
        import ctypes as C

        # every rank must be able to map rank 0's buffer (CUDA IPC + peer access); if any cannot, all fall back to the collective
        ok = 1
        handle = torch.zeros(64, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            sink_base = eng.lib.qb_device_alloc(eng.h, world * shard_bytes)
            hb = (C.c_ubyte * 64)()
            if not sink_base or eng.lib.qb_ipc_export(eng.h, sink_base, hb) != 0:
                ok = 0
            else:
                handle.copy_(torch.tensor(list(hb), dtype=torch.uint8))
        dist.broadcast(handle, src=0)
        if rank != 0:
            hb = (C.c_ubyte * 64)(*handle.cpu().tolist())
            ptr = C.c_void_p()
            if eng.lib.qb_ipc_open(eng.h, hb, C.byref(ptr)) != 0 or not ptr.value:
                ok = 0
            else:
                sink_base = ptr.value
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag[0]) == 0:
            if sink_base and rank != 0:
                eng.lib.qb_ipc_close(eng.h, sink_base)
            dist.barrier()
            if sink_base and rank == 0:
                eng.lib.qb_device_free(eng.h, sink_base)
            sink_base = None
            gather_mode = "nccl"
        else:
            sink_local = sink_base + rank * shard_bytes
            eng._check(eng.lib.qb_gblock_set_sink(eng.h, sink_local), "qb_gblock_set_sink")
    if world > 1 and gather_mode != "peer":
        mine = torch.as_tensor(_DevBytes(eng.lib.qb_gblock_device_tables(eng.h), shard_bytes), device=f"cuda:{local}")
        gathered = [torch.empty(shard_bytes, dtype=torch.uint8, device=f"cuda:{local}") for _ in range(world)] if rank == 0 else None

    gev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * (args.steps + args.warmup))]

    def step(i):
        # all kernels over the resident shard (front groups and chunks spread over the engine's streams, joined on the
        # handle's stream) -- with the peer sink the copies to rank 0 ride along; with QB_GATHER=nccl the collective follows
        # on the same stream (NCCL waits for the tables and the stream waits for NCCL, no host synchronisation)
        eng.run()
        if mine is not None:
            with torch.cuda.stream(stream):
                gev[2 * i].record(stream)
                dist.gather(mine, gathered, dst=0)
                gev[2 * i + 1].record(stream)

    for i in range(args.warmup):
        step(i)
    eng.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = eng.launches
    t_wall = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for i in range(args.steps):
        # every step rewrites its tables, series and scaled-coefficient arrays (tens of GB, far beyond the 126 MB L2)
        step(args.warmup + i)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches_timed = eng.launches - launches_before
    dev_ms = ev0.elapsed_time(ev1)
    gather_ms = 0.0
    if mine is not None:
        gather_ms = sum(gev[2 * (args.warmup + i)].elapsed_time(gev[2 * (args.warmup + i) + 1]) for i in range(args.steps)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    gather_ok, nogather_ms = None, 0.0
    if world > 1 and gather_mode == "peer":
        # (not timed) what rank 0 holds is what the ranks computed: a checksum of every rank's tables against the same
        # checksum of its slice of rank 0's buffer
        local_t = torch.as_tensor(_DevBytes(eng.lib.qb_gblock_device_tables(eng.h), shard_bytes), device=f"cuda:{local}")
        cs = local_t.view(torch.int32).to(torch.int64).sum().reshape(1)
        css = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(css, cs)
        if rank == 0:
            allt = torch.as_tensor(_DevBytes(sink_base, world * shard_bytes), device=f"cuda:{local}").view(world, -1)
            gather_ok = all(int(allt[r].view(torch.int32).to(torch.int64).sum()) == int(css[r]) for r in range(world))
        # and what the gather costs on top of the kernels: the same steps without the sink
        eng._check(eng.lib.qb_gblock_set_sink(eng.h, None), "qb_gblock_set_sink")
        eng.run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(args.steps):
            eng.run()
        with torch.cuda.stream(stream):
            e1.record(stream)
        barrier()
        nogather_ms = e0.elapsed_time(e1) / args.steps

    # per-kernel device times (CUDA events on the launching stream around every stage; chunks run back to back on one
    # stream in this mode so that the events bracket exactly one stage each)
    eng.set_timing(True)
    kms = np.zeros(4)
    ksteps = max(1, min(args.steps, 2))
    eng.run()
    eng.kernel_times()
    for _ in range(ksteps):
        eng.run()
        kms += np.array(eng.kernel_times())
    kms /= ksteps  # per step

    # ---- F/H blocks over the resident tables (Context::F_block / evaluate): terms per second through
    # qb_fblock_batch, results to the host (N = 1; inside a PMP build the same products are fused into k_pmp_accum) -------
    fblock = None
    if world == 1:
        import ctypes as C

        m = min(n, 32768)
        t_idx = np.ascontiguousarray(np.arange(m, dtype=np.int32) * (n // m))
        dvals = np.stack([qb.from_str("0.5181489", PREC), qb.from_str("0.96543695", PREC), qb.from_str("1.412625", PREC)])
        d_idx = np.ascontiguousarray(np.arange(m, dtype=np.int32) % 3)
        sym = np.ascontiguousarray(np.arange(m, dtype=np.int32) % 2)
        dims = [eng.lib.qb_block_dim(eng.h, 0), eng.lib.qb_block_dim(eng.h, 1)]
        f_out = np.zeros((int(sum(dims[int(s)] for s in sym)), W), dtype=np.uint32)
        P_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        for rep in range(3):
            t0 = time.perf_counter()
            eng._check(eng.lib.qb_fblock_batch(eng.h, m, P_(t_idx), len(dvals), P_(dvals), P_(d_idx), P_(sym), P_(f_out)), "qb_fblock_batch")
            t_f = time.perf_counter() - t0
        pairs = 2640  # Mixed x Mixed products per term at lambda = 19
        fblock = {"terms": int(m), "terms_per_s": m / t_f, "ms": 1e3 * t_f, "d2h_bytes": int(f_out.nbytes),
                  "limb_macs_per_term": pairs * 32 * 32, "frac_of_peak": m * pairs * 32 * 32 / t_f / peak,
                  "note": "wall clock of the C-ABI call: upload of the term list, k_vtod + k_fblock, results copied to pageable host memory"}
        del f_out

    # ---- end-to-end leg: host buffers in, host buffers out, through the one-shot C-ABI call ------------------------------
    lib, h = eng.lib, eng.h
    import ctypes as C

    def e2e_once():
        rc = lib.qb_gblock_batch(h, n, C.c_void_p(h_spin.data_ptr()), C.c_void_p(h_delta.data_ptr()), C.c_void_p(h_S.data_ptr()),
                                 C.c_void_p(h_P.data_ptr()), C.c_void_p(h_out.data_ptr()))
        if rc != 0:
            raise SystemExit(f"qb_gblock_batch failed: {rc}")

    eng.set_timing(False)
    e2e_once()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_once()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    # ---- reduce over ranks ------------------------------------------------------------------------------------------------
    if sink_base:
        # every rank has stopped writing (barriers above): unmap, then rank 0 frees
        if rank != 0:
            eng.lib.qb_ipc_close(eng.h, sink_base)
        dist.barrier()
        if rank == 0:
            eng.lib.qb_device_free(eng.h, sink_base)
    stats = torch.tensor([dev_ms, e2e_s, gather_ms, float(n), nogather_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, gather_ms, total_tables, nogather_ms = float(mx[0]), float(mx[1]), float(mx[2]), float(sm[3]), float(mx[4])
    else:
        total_tables = float(n)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_per_step = dev_ms / args.steps
    value = total_tables / (ms_per_step * 1e-3)
    # distinct keys of the step: the scan re-emits the even sector's keys (S = P = 0, Delta independent of the grid
    # point) at every grid point; the table kernels compute what they are handed -- the façade's planner is what dedupes
    unique_tables = None
    if world == 1:
        rows = np.ascontiguousarray(np.concatenate([spin.astype(np.uint32)[:, None], delta, S, P], axis=1))
        unique_tables = int(len(np.unique(rows.view(np.dtype((np.void, rows.dtype.itemsize * rows.shape[1]))))))
        del rows
    in_bytes = int(spin.nbytes + delta.nbytes + S.nbytes + P.nbytes)
    out_bytes = int(n * table_bytes)
    stage_names = ("rec", "series", "deriv", "finish")
    stages = {}
    for i, s in enumerate(stage_names):
        w = total_work(spin, s)
        stages[s] = {"ms_per_step": float(kms[i]), "limb_macs": float(w), "frac_of_peak": float(w / (kms[i] * 1e-3) / peak) if kms[i] > 0 else None}
    step_frac = total_work(spin) * world / (ms_per_step * 1e-3) / (peak * world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f1024 (32 x u32 limbs, round-to-nearest-even per operation)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "total_tables_per_step": int(total_tables), "tables_per_gpu_per_step": n,
                   "distinct_keys_per_step": unique_tables, "precision_bits": PREC,
                   "n_Max": N_MAX, "lambda": LAM, "sharding": "keys sorted by cost class (spin 0 / spin > 0), dealt round-robin to the ranks",
                   "l2": "each step rewrites its %d MB of tables and %d MB of series coefficients (> 126 MB L2)" % (
                       out_bytes // 2**20, n * (N_MAX + 1) * W * 4 // 2**20)},
        "e2e": {"value": total_tables / e2e_s, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes},
        "gpu_launches": int(launches_timed),
        "kernel_ms_per_step": {"rec_coeffs": float(kms[0]), "series": float(kms[1]), "deriv": float(kms[2]), "finish": float(kms[3]),
                               "note": "chunks serialised on one stream with an event around every stage; the timed step overlaps them"},
        "roofline": {"bound": "imad", "achieved": step_frac * peak / 1e12, "peak": peak / 1e12, "unit": "Tlimb-MAC/s", "frac": step_frac,
                     "traffic": ncu_traffic(n), "traffic_note": "DRAM bytes per step per GPU from the ncu capture of every kernel (profiles/r2_traffic.json); "
                                                                "the path is multiply-bound, not HBM-bound: ~7.2 MB per table against 2.3e7 limb-MACs",
                     "kernel": "whole step (k_rec_coeffs, k_series, k_pow, k_scale, k_horner, k_rho_to_z, k_offdiag_*)",
                     "stages": stages,
                     "peak_source": "measured on this GPU by qb_imad_peak: dependency-free IMAD.WIDE.U32 chains at full occupancy "
                                    "(32 x 32 + 64 -> 64; issues at 32 per clock per SM, half the 32-bit IMAD rate)",
                     "work": "algorithmic limb-MACs of the reference algorithm per table (SURVEY 8(d)): mul/fma L^2, div 2 L^2, pow 64 L^2"},
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if world > 1:
        nv = int((world - 1) * n * table_bytes)
        if gather_mode == "peer":
            line["gather"] = {"mode": "peer", "nvlink_bytes_per_step": nv, "to": "rank 0 HBM", "verified": gather_ok,
                              "how": "every finished chunk is copied by the copy engine into the rank's slice of rank 0's buffer (CUDA IPC, "
                                     "peer access over NVLink) underneath the kernels of the later chunks; the step ends when the copies have landed",
                              "ms_per_step_without_gather": nogather_ms, "exposed_ms_per_step": ms_per_step - nogather_ms}
        else:
            line["gather"] = {"mode": "nccl", "ms_per_step": gather_ms, "nvlink_bytes_per_step": nv, "to": "rank 0 HBM",
                              "how": "torch.distributed.gather over NCCL (grouped send/recv), enqueued behind the tables on the engine's stream",
                              "gbytes_per_s": (nv / 1e9) / (gather_ms * 1e-3) if gather_ms > 0 else None}
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import ref

            if ref.available():
                cores = host_cores()
                sample = max(64, 16 * cores)
                idx = strided_sample(n, sample)
                t = ref.time_gblock(PREC, N_MAX, LAM, DIM, delta[idx], spin[idx], S[idx], P[idx], cores)
                line["cpu_baseline"] = {"value": len(idx) / t, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{len(idx)} tables strided over the same workload (all sectors, spins and (S, P) classes) through the "
                                                  f"unmodified reference (qboot::gBlock, MPFR 4.2.1) on {cores} host threads"}
                # not timed: the tables the end-to-end leg just wrote to host memory, against the reference's, word for word
                idx = strided_sample(n, 256)
                want = ref.gblock(PREC, N_MAX, LAM, DIM, delta[idx], spin[idx], S[idx], P[idx])
                got = h_out.numpy().view(np.uint32)[idx]
                line["parity_check"] = {"tables": int(len(idx)), "bit_identical": bool(np.array_equal(got, want)),
                                        "against": "unmodified reference (oracle/_ref) on the same keys"}
        except Exception as ex:  # the baseline is reported, never required
            line["cpu_baseline"] = {"error": str(ex)}
    if fblock is not None:
        line["fblock"] = fblock
    if world == 1 and not args.no_pmp:
        # M2.  The table engine's buffers are released first: the facade owns its own handle
        del h_out
        eng.close()
        torch.cuda.empty_cache()
        try:
            from oracle import ref

            line["pmp_build"] = pmp_leg(host_cores(), ref.available() and not args.no_cpu_baseline)
        except Exception as ex:
            line["pmp_build"] = {"error": str(ex)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
