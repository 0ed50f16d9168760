#!/usr/bin/env python
"""bench.py — examples/sec of the DCN-v2 Criteo-shape training step (BASELINE.json configs[1], "C2":
26 categorical features, vocab 1e6 each, embed_dim 32, batch 65536, 3 full-rank cross layers,
Dense 192-192-1, MSE, AdamW — the examples/dcn.py wiring) on N B200s, plus the fused embedding-gather
HBM roofline and the CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = gather -> 3x FeatureCross -> MLP -> MSE -> full backward -> embedding scatter-add ->
AdamW on every parameter (the dense-gradient semantics of the reference: every table row is
visited).  `value` times K steps with the batch already in HBM; `e2e` times the same step driven
through the public API from PINNED HOST buffers (H2D of ids+labels and a D2H read of the loss inside
the timed region).  Inputs are far larger than L2 (3.3 GB of tables, random rows), so no flush is
needed between iterations ("l2": "inputs_larger_than_L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "adagrad", "sgd"])
    ap.add_argument("--engine", default="auto", choices=["auto", "ffma", "tcgen05", "tcgen05_ts"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--features", type=int, default=26)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--embed-dim", type=int, default=32)
    ap.add_argument("--cross-layers", type=int, default=3)
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-time bound of the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-variant", type=int, default=0)
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (1 GPU; the eager step already has no launch gaps)")
    ap.add_argument("--projection-dim", type=int, default=0, help="low-rank FeatureCross (ml_perf uses 512); 0 = full rank")
    ap.add_argument("--dense-units", default="192,192")
    return ap.parse_args()


def workload_name(a):
    rank_s = f"low-rank P={a.projection_dim}" if a.projection_dim else "full-rank"
    return (f"DCN-v2 C2: {a.features} categorical features, vocab {a.vocab} each, embed_dim={a.embed_dim}, "
            f"batch={a.batch}, {a.cross_layers} {rank_s} cross layers, Dense {a.dense_units.replace(',', '-')}-1, MSE, "
            f"{a.optimizer}")


# --------------------------------------------------------------------------------------- reference arm
def run_reference(a):
    """The reference's own CPU path for this workload.  keras/jax cannot be installed here,
    so this times the oracle PORT (oracle/torch_ref.py: the Keras op sequence one-for-one in torch-CPU),
    all host threads, each step one full batch of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_ref as T
    cores = T.usable_cores()        # min(cpu_count, affinity, cgroup quota): the threads the host really grants
    # a full step of this workload costs ~50 s of CPU time (the dense AdamW sweep over 3.3 GB of tables dominates),
    # so the run is bounded: 1 warm-up step, then timed steps until --cpu-budget-s is spent
    r = T.time_cpu_baseline(B=a.batch, F=a.features, V=a.vocab, E=a.embed_dim, L=a.cross_layers, steps=a.steps,
                            warmup=min(a.warmup, 1), optimizer=a.optimizer, threads=cores, budget_s=a.cpu_budget_s)
    line = {
        "impl": "reference", "metric": "examples/sec", "value": r["value"], "unit": "examples/s", "n_gpus": a.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "requested_steps": a.steps, "requested_warmup": a.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": a.batch},
        "cpu_baseline": {"value": r["value"], "unit": "examples/s", "cores": cores, "kind": "port",
                         "sample": f"{r['steps']} timed full step(s) of batch {a.batch} after {r['warmup']} warm-up, bounded to "
                                   f"{a.cpu_budget_s:.0f} s (torch-CPU restatement of the Keras op sequence on all host "
                                   "cores; keras/jax not installable)"},
        "e2e": {"value": r["value"], "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import keras_rs_b200 as K
    from keras_rs_b200._lib import check, lib, ptr, stream
    from keras_rs_b200.dcn import DCN

    engine = a.engine
    if engine == "auto":
        engine = os.environ.get("KRS_GEMM_ENGINE", "tcgen05_ts")   # tensor-pipe 3xTF32 (fp32-level accuracy); "ffma" = exact fp32 FMA
    K.set_gemm_engine(engine)

    B, F, V, E, L = a.batch, a.features, a.vocab, a.embed_dim, a.cross_layers
    units = tuple(int(u) for u in a.dense_units.split(",") if u)
    proj = a.projection_dim or None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    if world > 1:
        from keras_rs_b200.sharded import ShardedDCN
        model = ShardedDCN([V] * F, embedding_dim=E, num_cross_layers=L, dense_units=units, projection_dim=proj,
                           seed=1234, rank=rank, world=world)
    else:
        model = DCN([V] * F, embedding_dim=E, num_cross_layers=L, dense_units=units, projection_dim=proj, seed=1234)
    opt = {"adamw": lambda: K.optimizers.AdamW(0.01), "adagrad": lambda: K.optimizers.Adagrad(0.01),
           "sgd": lambda: K.optimizers.SGD(0.01)}[a.optimizer]()

    # synthetic Criteo-shaped data: NB distinct batches, pinned on the host and resident on the device
    NB = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host_ids = [torch.randint(0, V, (B, F), generator=g, dtype=torch.int32).pin_memory() for _ in range(NB)]
    host_y = [torch.rand((B,), generator=g).pin_memory() for _ in range(NB)]
    dev_ids = [t.cuda() for t in host_ids]
    dev_y = [t.cuda() for t in host_y]
    denom = B * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- device-resident steps ("value") --------------------------------------------------
    use_graph = a.graph and world == 1
    if use_graph:
        try:                                           # capture once (also validates NCCL capture on N > 1)
            model.train_on_batch_graph(dev_ids[0], dev_y[0], opt, denom)
            torch.cuda.synchronize()
        except Exception as exc:                       # pragma: no cover
            if rank == 0:
                print(f"# CUDA-graph capture unavailable ({type(exc).__name__}: {exc}); eager launches", file=sys.stderr)
            use_graph = False
    train = model.train_on_batch_graph if use_graph else model.train_on_batch

    def step_dev(i):
        train(dev_ids[i % NB], dev_y[i % NB], opt, denom)

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    split0 = lib.krs_gemm_split_launch_count()
    ms_total = timed(step_dev, a.steps, a.warmup)
    split_launches = (lib.krs_gemm_split_launch_count() - split0) * a.steps // (a.steps + a.warmup)   # timed steps only
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    value = B * world / (ms_step * 1e-3)

    # ---- end-to-end through the public API from pinned host buffers ("e2e") ----------------
    loss_host = torch.zeros((a.steps + a.warmup + 1,), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        loss = train(host_ids[i % NB], host_y[i % NB], opt, denom)
        loss_host[i % loss_host.numel():i % loss_host.numel() + 1].copy_(loss, non_blocking=True)

    ms_e2e = timed(step_e2e, a.steps, 3) / a.steps
    e2e_value = B * world / (ms_e2e * 1e-3)
    h2d = host_ids[0].numel() * 4 + host_y[0].numel() * 4
    final_loss = float(loss_host[(a.steps + 2) % loss_host.numel()])

    # ---- per-kernel measurements (rank 0 GPU, device resident, CUDA events on the launch stream) ----
    kern = {}
    base = model.local if hasattr(model, "local") else model
    if world == 1:
        bufs = base._step_buffers(B)
        plan = bufs["plan"]
        out = bufs["xs"][0]
        s = stream()
        # distinct id batches per launch so rows are re-fetched from HBM
        plans = []
        for i in range(NB):
            p = K.ops.GatherPlan(base._feature_list(dev_ids[i]))
            plans.append(p)

        def gather_i(i):
            p = plans[i % NB]
            check(lib.krs_gather_fwd(p.arr, p.F, B, ptr(out), base.D, a.gather_variant, s))

        reps = 20
        ms_g = timed(gather_i, reps, 5) / reps
        gather_bytes = B * F * E * 4 * 2 + B * F * 4          # rows read + out written + int32 ids
        kern["gather_fwd"] = {"ms": ms_g, "GBps": gather_bytes / ms_g * 1e-6, "bytes": gather_bytes}
        c0 = base.cross[0]
        x0, x1, h2 = bufs["xs"][0], bufs["xs"][1], bufs["h2"][0]
        hp0 = bufs["hproj"][0]

        def cross_i(i):
            check(lib.krs_cross_fwd(ptr(x0), ptr(x0), ptr(c0.down_proj_kernel), ptr(c0.kernel), ptr(c0.bias), 0.0, 0, ptr(x1),
                                    ptr(h2), None, ptr(hp0), B, base.D, proj or 0, s))

        ms_c = timed(cross_i, 5, 2) / 5
        flops = 4.0 * B * base.D * proj if proj else 2.0 * B * base.D * base.D
        kern["cross_fwd"] = {"ms": ms_c, "TFLOPs": flops / ms_c * 1e-9, "flops": flops, "engine": engine}

        def adamw_i(i):
            opt._update(base.emb, base.emb_grad, base.emb_touched)

        if a.optimizer == "adamw":     # dense-exact sweep: p, m, v read + written, g read + re-zeroed
            ms_a = timed(adamw_i, 5, 2) / 5
            # rows that ever received a gradient move 24 B per parameter (p, m, v read + written), the others 8 B (p only)
            ever = getattr(base, "emb_ever", None)
            hot = 1.0 if ever is None else float(torch.tensor([bin(int(x) & 0xFFFFFFFF).count("1") for x in ever[:65536].tolist()]).sum()) / (65536 * 32.0)
            adam_bytes = int(base.emb.numel() * 4 * (6 * hot + 2 * (1.0 - hot)))
            kern["adamw_tables"] = {"ms": ms_a, "GBps": adam_bytes / ms_a * 1e-6, "bytes": adam_bytes}

        def scatter_i(i):
            p = plans[i % NB]
            for f in range(F):
                p.arr[f].grad = base.emb_grad[base.row_off[f]:].data_ptr()
                p.arr[f].touched = base.emb_touched[base.row_off[f] // 32:].data_ptr()
            check(lib.krs_gather_bwd(p.arr, p.F, B, ptr(bufs["ga"]), base.D, s))

        ms_s = timed(scatter_i, 5, 2) / 5
        kern["gather_bwd"] = {"ms": ms_s, "GBps": (B * F * E * 4 * 3 + B * F * 4) / ms_s * 1e-6}
        base.emb_grad.zero_(); base.emb_touched.zero_()

    # ---- in-situ per-call profile of the eager step (CUDA events around every C-ABI call) -------------------
    step_profile = None
    if world == 1:
        import collections
        names = ["krs_gather_fwd", "krs_cross_fwd", "krs_dense_fwd", "krs_loss_fwd_bwd", "krs_dense_bwd", "krs_cross_bwd",
                 "krs_gather_bwd", "krs_adamw", "krs_sgd_adagrad", "krs_adam_hyper_advance"]
        recs = []
        orig = {n: getattr(lib, n) for n in names}

        def wrap(n, fn):
            def f(*args):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args)
                e1.record()
                recs.append((n, e0, e1))
                return rc
            return f

        for n in names:
            setattr(lib, n, wrap(n, orig[n]))
        try:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(2):
                base.train_on_batch(dev_ids[i % NB], dev_y[i % NB], opt, denom)
            recs.clear()
            t0.record()
            for i in range(3):
                base.train_on_batch(dev_ids[i % NB], dev_y[i % NB], opt, denom)
            t1.record()
            torch.cuda.synchronize()
            agg = collections.OrderedDict()
            for n, e0, e1 in recs:
                agg[n] = agg.get(n, 0.0) + e0.elapsed_time(e1) / 3
            step_profile = {k: round(v, 3) for k, v in agg.items()}
            step_profile["sum_of_calls"] = round(sum(agg.values()), 3)
            step_profile["step_wall"] = round(t0.elapsed_time(t1) / 3, 3)
        finally:
            for n in names:
                setattr(lib, n, orig[n])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on the host cores (bounded sample; rank 0, N=1 only) --------------------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        del model
        torch.cuda.empty_cache()
        from oracle import torch_ref as T
        r = T.time_cpu_baseline(B=B, F=F, V=V, E=E, L=L, steps=a.cpu_steps, warmup=1, optimizer=a.optimizer,
                                budget_s=a.cpu_budget_s)
        cpu = {"value": r["value"], "unit": "examples/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} timed full step(s) of batch {B} after 1 warm-up, bounded to {a.cpu_budget_s:.0f} s "
                         "(torch-CPU restatement of the Keras op sequence on all host cores; keras/jax not installable "
                         "here)", "ms_per_step": r["ms_per_step"]}

    n_mlp = len(units) + 1
    launches_per_step = 1 + L * (2 if proj else 1) + n_mlp + 1 + 3 * n_mlp + L * (5 if proj else 3) + 1 + 2
    # ncu-measured DRAM traffic per launch of the two kernels below (profiles/ncu_traffic.json, written from the
    # committed `ncu --set full` captures by profiles/summarize.py; null when no capture of that kernel is committed)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    roof = roof_gather = None
    if kern:
        gk = kern["gather_fwd"]
        roof_gather = {"kernel": "gather_fast_kernel (fused 26-table gather+concat)", "bound": "hbm", "achieved": gk["GBps"],
                       "peak": hbm_peak, "unit": "GB/s", "frac": gk["GBps"] / hbm_peak,
                       "frac_of_8TBps_nominal": gk["GBps"] / 8000.0, "peak_source": peak_src,
                       "traffic": (traffic.get("gather_fast_kernel") or {}).get("dram_bytes"),
                       "algorithmic_bytes": gk["bytes"], "ms": gk["ms"]}
        # the dominant kernel of the step is the dense contraction (gemm_tc_kernel: ~60 % of the step's device time,
        # profiles/r1_launches_tcgen05.md).  Algorithmic flops = 2*B*D*D per full-rank cross layer; the
        # kernel issues 3 TF32 MMAs per product for fp32-level accuracy, so its tensor peak is the measured dense bf16
        # rate / 2 (tf32) / 3 (split) = bf16 / 6 — the burst figure, since this launch is timed alone.
        ck = kern["cross_fwd"]
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        tpeak = bf16_peak / 6.0 if engine != "ffma" else None
        roof = {"kernel": ("gemm_tc_kernel" if engine != "ffma" else "sgemm_kernel") + " (FeatureCross forward: x@W + fused cross epilogue)",
                "bound": "tensor", "achieved": ck["TFLOPs"], "peak": tpeak, "unit": "TFLOP/s",
                "frac": (ck["TFLOPs"] / tpeak) if tpeak else None,
                "peak_source": ("measured bf16 burst %.0f TF/s (MEASURED_PEAKS.json) / 6 for the 3xTF32 split" % bf16_peak)
                if "bf16_tflops" in peaks else "fallback 1590 TF/s bf16 / 6",
                "algorithmic_flops": ck["flops"], "ms": ck["ms"], "engine": engine,
                "traffic": (traffic.get("gemm_tc_kernel") or {}).get("dram_bytes")}
    line = {
        "metric": "examples/sec", "value": value, "unit": "examples/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": B * world, "parallelism": f"dp{world}" + (
            "+mod-row-sharded tables over NVLink peer memory" if world > 1 else ""), "gemm_engine": engine,
            "l2": "inputs_larger_than_L2", "final_loss": final_loss, "launch": "cuda_graph" if use_graph else "eager"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "examples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        # kernels executed in the timed region: the fixed launch sequence of the step x steps, plus the B_lo plane
        # kernels the tcgen05 engines launched in front of weight GEMMs (counted by the library)
        "gpu_launches": launches_per_step * a.steps + split_launches,
        "roofline": roof, "roofline_gather": roof_gather, "kernels": kern, "step_profile_ms": step_profile, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
