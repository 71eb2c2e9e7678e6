#!/usr/bin/env python
"""bench.py — train slides/sec on 16k-patch bags (BASELINE.json metric), one process per GPU.

Workload (BASELINE.json configs[1]): Nystrom-attention TransMIL (dim 512, 8 heads x 64, 256 landmarks, 6 pinv
iterations, PPEG, 2 layers) on synthetic bags of 16 384 x 1024-d patch embeddings; one "step" = forward + CE loss +
backward + Adam step on `--batch` slides per GPU (data parallel over slides, gradient all-reduce over NCCL).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host cores

Prints ONE JSON line (rank 0).
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
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

METRIC = "train slides/sec at 16k-patch bags"
UNIT = "slides/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SML_PRECISION", "bf16"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--batch", type=int, default=8, help="slides per step per GPU (config_mine.yaml batch_size: 8)")
    ap.add_argument("--patches", type=int, default=16384)
    ap.add_argument("--workload", default="transmil", choices=["transmil", "deform2d", "deform1d", "pathomic"],
                    help="transmil = BASELINE.json configs[1] (the headline); deform2d / deform1d = configs[2], the full "
                         "DeformCrossTransMIL (attn_dim 2 / 1) train step; pathomic = DeformPathomicNet (two genomic encoders + two "
                         "DeformCrossTransMIL branches + heads) with the reference's loss: weighted CE + BatchLoss over all ranks "
                         "(train_test.py:63-83)")
    ap.add_argument("--cpu-sample", type=int, default=5, help="slides in the CPU baseline sample (~2 s each on 16 host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prof-out", default=None, help="write the per-shape GEMM timing table of the instrumented pass here")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md clocks line), sampled every 50 ms through NVML
    (no process spawn; nvidia-smi is the fallback).  Rank 0 samples its own GPU only: eight samplers on one host would
    perturb what they measure."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(v) for v in vis.split(",") if v.strip() != ""]
            index = ids[index] if ids else index
        except ValueError:
            pass
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.handle))
        act = lambda bit: "Active" if r & bit else "Not Active"
        return [str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)]     # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [s.strip() for s in out.strip().split(",")]
                    if len(f) >= 6:
                        self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def synth_bag(batch, patches, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(batch, patches, 1024, generator=g, device=device)
    y = torch.randint(0, 4, (batch,), generator=g, device=device)
    return x, y


# ------------------------------------------------------------------------------------------- reference arm / CPU
def cpu_port_step_fn(patches, threads):
    """The reference algorithm (oracle restatement of models/mil.py:209-259 + NystromAttention) on host cores."""
    import oracle
    from argparse import Namespace
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    import sml_b200
    shapes = {k: tuple(v.shape) for k, v in sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).state_dict().items()}
    from synth import synth_state_dict
    params = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in synth_state_dict(shapes, 7).items()}
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3, weight_decay=0.1)

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        _, logits = oracle.nystrom.transmil(x, params)
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step = cpu_port_step_fn(args.patches, threads)
    x, y = synth_bag(1, args.patches, 42, "cpu")
    for _ in range(min(args.warmup, 1)):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(x, y)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TransMIL(Nystrom m=256, 6 pinv iters, PPEG, 2 layers) fwd+CE+bwd+Adam on %d-patch x 1024-d bags"
                                   % args.patches, "slides_per_step_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
                       "precision": "fp32", "parallelism": "host threads",
                       "arm": "oracle port of the reference algorithm (models/mil.py + NystromAttention) on the host cores; each "
                              "timed step is a bounded sample of 1 slide of the workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d steps x 1 slide of %d patches, fp32, %d of %d host threads"
                                       % (args.steps, args.patches, torch.get_num_threads(), threads)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from argparse import Namespace
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = None
    if world > 1:
        # NCCL (NCCL_DEBUG=VERSION in this image) writes its version banner to fd 1; rank 0 has to print ONE JSON line, so
        # everything else goes to stderr and the line is written to the saved descriptor at the end
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    import sml_b200
    from sml_b200 import functional as Fn
    from sml_b200 import ops

    torch.manual_seed(42)
    deform = args.workload != "transmil"
    pathomic = args.workload == "pathomic"
    if pathomic:
        margs = Namespace(init_type="max", input_size_omic_tumor=59, input_size_omic_immune=361, omic_dim=128, path_dim=128,
                          mmhid=128, return_grad="False", dropout_rate=0.1, label_dim=4, fusion_type="concat", cut_fuse_grad=False,
                          return_vgrid=True, attn_dim=2, task_type="diag2021")          # config/config_mine.yaml
        model = sml_b200.DeformPathomicNet(margs).to(dev).train()
        model.set_precision(args.precision)
        bloss = sml_b200.BatchLoss(args.batch, world)
        ce_w = torch.tensor([1.0, 4.15, 2.93, 2.43], device=dev)                        # train_test.py:25
    elif deform:
        model = sml_b200.DeformCrossTransMIL(Namespace(path_dim=128, attn_dim=2 if args.workload == "deform2d" else 1,
                                                       return_vgrid=False)).to(dev).train()
        model.set_precision(args.precision)
    else:
        model = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).train()
        model.set_precision(args.precision)
    Fn.manual_seed(42 + rank)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=deform,  # DeformCrossTransMIL: the attn1d / attn2d twin module is unused
                                                        
                                                        broadcast_buffers=False, gradient_as_bucket_view=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.1, fused=True)
    B, N = args.batch, args.patches
    x_dev, y_dev = synth_bag(B, N, 42 + rank, dev)
    x_host = x_dev.cpu().pin_memory()
    y_host = y_dev.cpu().pin_memory()
    omic = torch.randn(B, 128, generator=torch.Generator(device=dev).manual_seed(7 + rank), device=dev).abs() if deform else None

    if pathomic:
        gen = torch.Generator(device=dev).manual_seed(9 + rank)
        x_t = torch.randn(B, 59, generator=gen, device=dev).abs()
        x_i = torch.randn(B, 361, generator=gen, device=dev).abs()

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        if pathomic:
            _, f_t, f_i, lg, _, _, _ = net(x_path=x, x_omic_tumor=x_t, x_omic_immune=x_i)
            loss = torch.nn.functional.cross_entropy(lg[2], y, weight=ce_w) + 0.5 * bloss(lg[3], lg[4]).sum() + 0.5 * bloss(lg[5], lg[6]).sum()
            loss.backward()
            sml_b200.modulate_classifier_grads(model.classifier, f_t, f_i, y)       # train_test.py:87-184 (gradient_modulate: True)
            opt.step()
            return loss
        if deform:
            logits = net(x, omic)[1]
            loss = torch.nn.functional.cross_entropy(logits, y)
        else:
            enc, logits, _ = net(x)
            # the multimodal_projection head takes part with zero weight so that every parameter gets a gradient and DDP needs
            # no unused-parameter search (the reference's loop feeds both heads, train_test.py:347-373)
            loss = torch.nn.functional.cross_entropy(logits, y) + 0.0 * enc.sum()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # samples from the last warm-up steps through the timed and e2e regions (all under load)
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    l0 = sml_b200._lib.launch_count()
    ms = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = sml_b200._lib.launch_count() - l0

    # ---- end to end: every step's bag comes from pinned host memory (H2D inside the timed region) and the loss is read
    #      back (D2H).  Like any loader, the copy of step i+1 is issued on a side stream while step i computes.
    copy_stream = torch.cuda.Stream(device=dev)
    xb = [torch.empty_like(x_dev) for _ in range(2)]
    yb = [torch.empty_like(y_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0}

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            xb[slot].copy_(x_host, non_blocking=True)
            yb[slot].copy_(y_host, non_blocking=True)
            ready[slot].record(copy_stream)

    for sl in range(2):
        freed[sl].record(torch.cuda.current_stream())
    issue_copy(0)

    def e2e_step():
        slot = state["i"] & 1
        state["i"] += 1
        issue_copy(slot ^ 1)                                   # prefetch the next step's bag
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = step(xb[slot], yb[slot])
        freed[slot].record(torch.cuda.current_stream())
        return float(loss.detach())                            # D2H read of the loss
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag = True

    # ---- per-kernel roofline: instrumented pass (CUDA events around every GEMM launch on the launch stream)
    ops.PROFILE = []
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    shapes = {}
    for eng, flops, e0, e1, tag in prof:
        dt = e0.elapsed_time(e1) * 1e-3
        a = agg.setdefault(eng, [0.0, 0.0, 0])
        a[0] += flops; a[1] += dt; a[2] += 1
        sh = shapes.setdefault((eng,) + tag, [0.0, 0.0, 0])
        sh[0] += flops; sh[1] += dt; sh[2] += 1
    if args.prof_out and rank == 0:
        with open(args.prof_out, "w") as f:
            f.write("# eng M N K batch A_MNmajor B_MNmajor split relu|2softmax C Cb_hi Cb_lo beta : launches total_us avg_us TFLOP/s\n")
            for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
                f.write("%s : %d %.1f %.1f %.1f\n" % (" ".join(str(x) for x in k), v[2], v[1] * 1e6, v[1] * 1e6 / v[2], v[0] / v[1] / 1e12))
    pk, pk_src = peaks()
    # per-shape roofline of the heaviest tcgen05 GEMM shapes: bound by their outputs (HBM) when K = 64, by the tensor pipe
    # otherwise.  `traffic` = dram read + write per launch from the round's `ncu --set full` capture (profiles/r01_ncu_summary.txt)
    ncu_traffic = {(256, 16640, 64, 64, 0): 138.8e6 + 1032.6e6, (16640, 256, 64, 64, 2): 138.6e6 + 492.2e6,
                   (133120, 1536, 512, 1, 0): 138.0e6 + 1170.8e6, (256, 256, 256, 64, 0): 33.6e6 + 0.9e6}
    by_shape = []
    for key, v in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
        if key[0] != "tc" or len(by_shape) >= 5:
            continue
        M_, N_, K_, nb_, _, _, split_, relu_, c_, hi_, lo_, _ = key[1:]
        t_avg = v[1] / v[2]
        bytes_ = nb_ * (2.0 * (1 + split_) * (M_ * K_ + K_ * N_) + M_ * N_ * (4.0 * c_ + 2.0 * hi_ + 2.0 * lo_))
        hbm_frac = bytes_ / t_avg / 1e9 / pk["hbm_gbs"]
        tc_frac = v[0] / v[1] / 1e12 / pk["bf16_tflops_sustained"]
        by_shape.append({"shape": "%dx%dx%d x%d%s%s" % (M_, N_, K_, nb_, " bf16x3" if split_ else "", " +softmax" if relu_ & 2 else ""),
                         "launches": v[2], "avg_us": t_avg * 1e6,
                         "bound": "latency (dependent chain, L2-resident operands)" if (K_ == M_ == N_ == 256) else
                                  ("hbm" if hbm_frac > tc_frac else "tensor"),
                         "achieved": (bytes_ / t_avg / 1e9) if hbm_frac > tc_frac else v[0] / v[1] / 1e12,
                         "unit": "GB/s" if hbm_frac > tc_frac else "TFLOP/s", "frac": max(hbm_frac, tc_frac),
                         "traffic": ncu_traffic.get((M_, N_, K_, nb_, relu_ & 2))})
    roof = None
    if deform and "deform_bwd" in agg:
        f, t, n = agg["deform_bwd"]
        ach = f / t / 1e12
        peak = pk["bf16_tflops_sustained"]
        roof = {"kernel": "deform_attn_bwd_mma_kernel (fused bias-MLP + attention backward, %d launch)" % n, "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": pk_src + " (sustained, tcgen05 dense bf16)", "share_of_step": t / (ms * 1e-3 / args.steps),
                "forward_kernel": {"achieved": agg["deform_fwd"][0] / agg["deform_fwd"][1] / 1e12,
                                   "share_of_step": agg["deform_fwd"][1] / (ms * 1e-3 / args.steps)},
                "note": "algorithmic FLOPs: 2*(2240 (2-D) | 2176 (1-D) + 256) per (query, key) pair for the backward; the in-kernel "
                        "forward recomputation and the 3-pass split are not counted.  Warp-level mma.sync (legacy tensor pipe, "
                        "~576 TFLOP/s), bounded by the per-pair CUDA-core work, not by the tensor pipe"}
    elif "tc" in agg and agg["tc"][1] > 0:
        f, t, n = agg["tc"]
        ach = f / t / 1e12
        peak = pk["bf16_tflops_sustained"]
        roof = {"kernel": "gemm_tc_kernel (tcgen05, all %d launches of one step)" % n, "bound": "tensor", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": pk_src + " (sustained)",
                "share_of_step": t / (ms * 1e-3 / args.steps),
                "note": "bf16x3 launches count 1x algorithmic FLOPs (3 MMA passes executed)", "by_shape": by_shape}
    elif "simt" in agg:
        f, t, n = agg["simt"]
        roof = {"kernel": "gemm_f32_kernel (SIMT fp32, %d launches)" % n, "bound": "tensor", "achieved": f / t / 1e12,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": f / t / 1e12 / pk["bf16_tflops_sustained"],
                "traffic": None, "peak_source": pk_src}

    if rank == 0:
        total = B * world * args.steps
        line = {"metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "f32 (bf16x3 split)", "fp32": "f32"}[args.precision],
                "data": "synthetic",
                "config": {"workload": ("DeformPathomicNet (2 x MaxNet + 2 x DeformCrossTransMIL attn_dim=2 + concat heads) fwd + weighted CE + "
                                        "BatchLoss(all ranks) + bwd + Adam on %d-patch x 1024-d bags" % N) if pathomic else
                                       ("DeformCrossTransMIL(attn_dim=%d: fc1 + genomic fusion + deformable cross-attention + heads) "
                                        "fwd+CE+bwd+Adam on %d-patch x 1024-d bags" % (2 if args.workload == "deform2d" else 1, N)) if deform
                           else "TransMIL(Nystrom m=256, 6 pinv iters, PPEG, 2 layers) fwd+CE+bwd+Adam on %d-patch x 1024-d bags"
                                       % N, "slides_per_step_per_gpu": B, "global_batch": B * world, "precision": args.precision,
                           "parallelism": "dp%d" % world, "l2": "inputs (%d MB/step) and activations exceed the 126 MB L2"
                                                              % (B * N * 1024 * 4 // 2 ** 20)},
                "gpu_launches": int(launches),
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 8),
                        "d2h_bytes_per_step": 4},
                "clocks": sampler.summary(), "roofline": roof}
        if not args.no_cpu_baseline and world == 1 and not deform:
            threads = os.cpu_count() or 1
            cstep = cpu_port_step_fn(N, threads)
            xs, ys = synth_bag(1, N, 42, "cpu")
            t0 = time.perf_counter()
            for _ in range(args.cpu_sample):
                cstep(xs, ys)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": args.cpu_sample / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "%d slide(s) of %d patches, fwd+CE+bwd+Adam, fp32 oracle port, %d host threads"
                                              % (args.cpu_sample, N, torch.get_num_threads())}
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
