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
    ap.add_argument("--workload", default="transmil", choices=["transmil", "deform2d", "deform1d", "pathomic", "pathomic1d"],
                    help="transmil = BASELINE.json configs[1] (the headline); deform2d / deform1d = configs[2], the full "
                         "DeformCrossTransMIL (attn_dim 2 / 1) train step; pathomic = DeformPathomicNet (two genomic encoders + two "
                         "DeformCrossTransMIL branches + heads) with the reference's loss: weighted CE + BatchLoss over all ranks "
                         "(train_test.py:63-83)")
    ap.add_argument("--cpu-sample", type=int, default=5, help="slides in the CPU baseline sample (~2 s each on 16 host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prof-out", default=None, help="write the per-shape GEMM timing table of the instrumented pass here")
    ap.add_argument("--secondary", default=os.environ.get("SML_BENCH_SECONDARY", "transmil_bf16x3,pathomic,pathomic1d,batchloss64,stress100k"),
                    help="comma list of extra workloads measured after the headline and reported under `secondary` "
                         "(transmil_bf16x3 = the fp32-grade mode of configs[1]; pathomic = configs[2]; batchloss64 = configs[3]: "
                         "BatchLoss + cross-rank gather + sync-BN BilinearFusion at global batch 64); empty string = none")
    ap.add_argument("--secondary-steps", type=int, default=5)
    ap.add_argument("--graph", type=int, default=int(os.environ.get("SML_BENCH_GRAPH", "1")),
                    help="1: replay the train step as one CUDA graph (sml_b200.GraphedStep) on single-GPU runs")
    return ap.parse_args()


def workload_name(workload, patches):
    return {"transmil": "TransMIL(Nystrom m=256, 6 pinv iters, PPEG, 2 layers) fwd+CE+bwd+Adam on %d-patch x 1024-d bags" % patches,
            "pathomic": "DeformPathomicNet (2 x MaxNet + 2 x DeformCrossTransMIL attn_dim=2 + concat heads) fwd + weighted CE + "
                        "BatchLoss(all ranks) + bwd + Adam on %d-patch x 1024-d bags" % patches,
            "pathomic1d": "DeformPathomicNet (2 x MaxNet + 2 x DeformCrossTransMIL attn_dim=1 [DeformableAttention1D] + concat heads) "
                          "fwd + weighted CE + bwd + classifier gradient modulation + Adam",
            "deform2d": "DeformCrossTransMIL(attn_dim=2: fc1 + genomic fusion + deformable cross-attention + heads) fwd+CE+bwd+Adam "
                        "on %d-patch x 1024-d bags" % patches,
            "deform1d": "DeformCrossTransMIL(attn_dim=1: fc1 + genomic fusion + deformable cross-attention + heads) fwd+CE+bwd+Adam "
                        "on %d-patch x 1024-d bags" % patches}[workload]


def workload_config(args, world):
    """the `config` object of the JSON line - identical for both arms (`--impl ours` / `--impl reference`)"""
    B, N = args.batch, args.patches
    return {"workload": workload_name(args.workload, N), "slides_per_step_per_gpu": B, "global_batch": B * world,
            "precision": args.precision, "parallelism": "dp%d" % world,
            "l2": "inputs (%d MB/step as bf16-packed bags, twice that in fp32) and activations exceed the 126 MB L2"
                  % (B * N * 1024 * 2 // 2 ** 20)}


def transmil_shapes():
    """parameter shapes of TransMIL (models/mil.py:209-259) from the committed golden fixture - the CPU arm must not import
    the product package (no libsml_b200.so in the reference process)"""
    return torch.load(os.path.join(ROOT, "tests", "golden", "transmil.pt"), weights_only=False)["shapes"]


PARAM_SEED = 7          # both arms start from synth_state_dict(transmil_shapes(), PARAM_SEED)


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
    """The reference algorithm (oracle restatement of models/mil.py:209-259 + NystromAttention) on host cores.  Imports only
    `oracle/` and the fixture helpers: nothing of the product package is loaded in this process."""
    import oracle
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    from synth import synth_state_dict
    params = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in synth_state_dict(transmil_shapes(), PARAM_SEED).items()}
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
    """`--impl reference`: the reference's CPU path (oracle port; the reference itself is 30 plain .py files with no
    setup.py, its TransMIL needs CUDA-only calls and an un-vendored PyPI package) on all host threads, on the
    SAME config / metric / unit / steps / warm-up as our arm.  Each step is a bounded sample of the configured step: ONE slide
    of the `--batch` slides (a full 8-slide step is ~13 s of CPU); value = slides / second."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    step = cpu_port_step_fn(args.patches, threads)
    x, y = synth_bag(1, args.patches, 42, "cpu")
    for _ in range(args.warmup):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(x, y)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "each of the %d timed steps = 1 slide (of the %d per configured step) of %d patches, "
                                       "fwd+CE+bwd+Adam, fp32 oracle port of models/mil.py + NystromAttention, %d of %d host threads"
                                       % (args.steps, args.batch, args.patches, torch.get_num_threads(), threads)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def load_ncu_traffic():
    """dram read + write bytes per launch of the dominant GEMM shapes, from the round's `ncu --set full` capture.  The file
    is written by tools/ncu_traffic.py from the .ncu-rep and names the capture it came from (commit, date); None when absent."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            d = json.load(open(path))
            return {tuple(e["key"]): e["bytes"] for e in d.get("shapes", [])}, {k: d.get(k) for k in ("source", "commit", "captured")}
    return {}, None


class Workload:
    """one train-step workload on this rank: model + optimizer + synthetic inputs + step function"""

    def __init__(self, name, precision, batch, patches, dev, rank, world, local, graph=False):
        import sml_b200
        from argparse import Namespace
        from sml_b200 import functional as Fn
        self.name, self.precision, self.B, self.N = name, precision, batch, patches
        self.deform = name in ("deform2d", "deform1d", "pathomic", "pathomic1d")
        torch.manual_seed(42)
        B, N = batch, patches
        if name in ("pathomic", "pathomic1d"):
            # pathomic1d = BASELINE configs[2] as written (DeformableAttention1D): attn_dim=1 cannot return vgrid in the reference
            # (SURVEY F11), so that configuration trains on the weighted CE alone; pathomic = the YAML default (attn_dim=2 + BatchLoss)
            one_d = name == "pathomic1d"
            margs = Namespace(init_type="max", input_size_omic_tumor=59, input_size_omic_immune=361, omic_dim=128, path_dim=128,
                              mmhid=128, return_grad="False", dropout_rate=0.1, label_dim=4, fusion_type="concat", cut_fuse_grad=False,
                              return_vgrid=not one_d, attn_dim=1 if one_d else 2, task_type="diag2021")   # config/config_mine.yaml
            model = sml_b200.DeformPathomicNet(margs).to(dev).train()
            self.bloss = None if one_d else sml_b200.BatchLoss(B, world)
            self.ce_w = torch.tensor([1.0, 4.15, 2.93, 2.43], device=dev)                   # train_test.py:25
            gen = torch.Generator(device=dev).manual_seed(9 + rank)
            self.x_t = torch.randn(B, 59, generator=gen, device=dev).abs()
            self.x_i = torch.randn(B, 361, generator=gen, device=dev).abs()
        elif self.deform:
            model = sml_b200.DeformCrossTransMIL(Namespace(path_dim=128, attn_dim=2 if name == "deform2d" else 1,
                                                           return_vgrid=False)).to(dev).train()
            self.omic = torch.randn(B, 128, generator=torch.Generator(device=dev).manual_seed(7 + rank), device=dev).abs()
        elif name == "stress100k":
            # configs[4]: TransMIL inference (the inference.py path: eval, no grad) on one 100 489-patch bag (317^2), 512 landmarks
            model = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev).eval()
            for lyr in (model.layer1, model.layer2):
                lyr.attn.num_landmarks = 512
        elif name == "batchloss64":
            # configs[3]: the subspace loss with its cross-rank gather (utils/loss.py:13-40, utils/gather.py) + the sync-BN
            # statistics exchange of BilinearFusion's two BatchNorm1d (models/fusion.py:29-32), global batch 64
            model = sml_b200.BilinearFusion(skip=0, use_bilinear=1, gate1=1, gate2=1, dim1=128, dim2=128, scale_dim1=1, scale_dim2=1,
                                            mmhid=128, dropout_rate=0.1).to(dev).train()
            self.bloss = sml_b200.BatchLoss(B, world)
            gen = torch.Generator(device=dev).manual_seed(11 + rank)
            self.vec1 = torch.randn(B, 128, generator=gen, device=dev)
            self.vec2 = torch.randn(B, 128, generator=gen, device=dev)
            self.vg = (torch.randn(8 * B, 2, 32, 32, generator=gen, device=dev) * 4 + 16).requires_grad_()
        else:
            from synth import synth_state_dict
            model = sml_b200.TransMIL(Namespace(label_dim=4, path_dim=128)).to(dev)
            model.load_state_dict(synth_state_dict(transmil_shapes(), PARAM_SEED))          # same start as the CPU arm
            model.train()
        if hasattr(model, "set_precision"):
            model.set_precision(precision)
        Fn.manual_seed(42 + rank)
        self.model = self.net = model
        self.gdp = bool(graph) and world > 1 and name == "transmil"   # graph replay + one flat gradient all-reduce instead of DDP
        if world > 1 and name != "stress100k" and not self.gdp:
            # DeformCrossTransMIL: the attn1d / attn2d twin module is unused -> find_unused_parameters
            self.net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=self.deform,
                                                                 broadcast_buffers=False, gradient_as_bucket_view=True)
        if name == "stress100k":
            self.net = model                                  # replicas only: no collective on this path
        self.graphed = None
        self.want_graph = bool(graph) and world == 1 and name in ("transmil", "pathomic", "pathomic1d", "deform2d", "deform1d")
        self.opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.1, fused=True, capturable=self.want_graph)
        if name == "batchloss64":
            self.x_dev = torch.zeros(1, device=dev); self.y_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        elif name == "stress100k":
            self.x_dev = torch.randn(1, 317 * 317, 1024, generator=torch.Generator(device=dev).manual_seed(42 + rank), device=dev)
            self.x_dev = self.x_dev.to(torch.bfloat16) if precision == "bf16" else self.x_dev
            self.y_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        else:
            self.x_dev, self.y_dev = synth_bag(B, N, 42 + rank, dev)
            self.x0_f32 = self.x_dev[:1].cpu()                     # fp32 original of slide 0: input of the CPU oracle probe
            if precision == "bf16":
                # bf16-packed bags (sml_b200.data.bags: the feed format of this mode).  Rounding the bag to bf16 is what the
                # mode's tensor-core GEMM does to its operand anyway, so the step computes bit-identical results from either
                self.x_dev = self.x_dev.to(torch.bfloat16)

    def enable_graph(self):
        """capture the step (after the eager warm-up) as one CUDA graph; on failure the workload stays eager and says so"""
        import sml_b200
        if self.gdp and self.graphed is None:
            F = torch.nn.functional

            def loss_fn(x, y):
                enc, logits, _ = self.model(x)
                return F.cross_entropy(logits, y) + 0.0 * enc.sum()
            self.graphed = sml_b200.GraphedDataParallelStep(self.model, self.opt, loss_fn, (self.x_dev, self.y_dev))
            self.graph_note = "CUDA graph of forward + backward into a flat gradient buffer, one NCCL all-reduce, fused Adam"
            return
        if not self.want_graph or self.graphed is not None:
            return
        try:
            self.graphed = sml_b200.GraphedStep(self._step, (self.x_dev, self.y_dev), warmup=1)
            self.graph_note = "one CUDA graph per step (sml_b200.GraphedStep)"
        except Exception as e:                               # noqa: BLE001
            self.graphed = None
            self.want_graph = False
            self.graph_note = "eager (graph capture failed: %s)" % str(e)[:120]
            import traceback
            print("graph capture of %s failed:\n%s" % (self.name, traceback.format_exc()), file=sys.stderr)

    def step(self, x, y):
        if self.graphed is not None:
            return self.graphed(x, y)
        return self._step(x, y)

    def _step(self, x, y):
        import sml_b200
        F = torch.nn.functional
        if self.gdp:                                         # eager form of the same step (warm-up, launch counting)
            self.opt.zero_grad(set_to_none=False)
            enc, logits, _ = self.model(x)
            loss = F.cross_entropy(logits, y) + 0.0 * enc.sum()
            loss.backward()
            import torch.distributed as dist
            for p_ in self.model.parameters():
                if p_.grad is not None:
                    dist.all_reduce(p_.grad, op=dist.ReduceOp.AVG)
            self.opt.step()
            return loss
        self.opt.zero_grad(set_to_none=True)
        if self.name in ("pathomic", "pathomic1d"):
            _, f_t, f_i, lg, _, _, _ = self.net(x_path=x, x_omic_tumor=self.x_t, x_omic_immune=self.x_i)
            loss = F.cross_entropy(lg[2], y, weight=self.ce_w)
            if self.bloss is not None:
                loss = loss + 0.5 * self.bloss(lg[3], lg[4]).sum() + 0.5 * self.bloss(lg[5], lg[6]).sum()
            loss.backward()
            sml_b200.modulate_classifier_grads(self.model.classifier, f_t, f_i, y)       # train_test.py:87-184 (gradient_modulate: True)
            self.opt.step()
            return loss
        if self.name == "stress100k":
            with torch.no_grad():
                return self.net(x)[1].sum()
        if self.name == "batchloss64":
            self.vg.grad = None
            emb = self.net(self.vec1, self.vec2)                                         # BilinearFusion: 2 x sync-BN (train mode)
            loss = self.bloss(emb, self.vg).sum() + emb.square().mean()                  # gather of [B,128] + vgrid, Gram epilogue
            loss.backward()
            self.opt.step()
            return loss
        if self.deform:
            logits = self.net(x, self.omic)[1]
            loss = F.cross_entropy(logits, y)
        else:
            enc, logits, _ = self.net(x)
            # the multimodal_projection head takes part with zero weight so that every parameter gets a gradient and DDP needs
            # no unused-parameter search (the reference's loop feeds both heads, train_test.py:347-373)
            loss = F.cross_entropy(logits, y) + 0.0 * enc.sum()
        loss.backward()
        self.opt.step()
        return loss


def run_ours(args):
    import torch.distributed as dist
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
    from sml_b200 import ops

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

    pk, pk_src = peaks()
    B, N = args.batch, args.patches
    wl = Workload(args.workload, args.precision, B, N, dev, rank, world, local, graph=args.graph)
    deform = wl.deform

    # ---- step-0 parity probe (rank 0): eval-mode logits / CE loss of slide 0 from the SAME initial parameters as the CPU
    #      arm, against the CPU oracle held in this process for the cpu_baseline leg
    parity = None
    if args.workload == "transmil" and rank == 0 and not args.no_cpu_baseline:
        import oracle
        from synth import synth_state_dict
        wl.model.eval()
        with torch.no_grad():
            _, lg, _ = wl.model(wl.x_dev[:1])
            loss_gpu = float(torch.nn.functional.cross_entropy(lg, wl.y_dev[:1]))
            p_cpu = synth_state_dict(transmil_shapes(), PARAM_SEED)
            _, lg_o = oracle.nystrom.transmil(wl.x0_f32, p_cpu)
            loss_o = float(torch.nn.functional.cross_entropy(lg_o, wl.y_dev[:1].cpu()))
        wl.model.train()
        tol = {"fp32": 1e-4, "bf16x3": 1e-4, "bf16": 2e-2}[args.precision]
        err_l = float((lg.cpu() - lg_o).abs().max() / lg_o.abs().max())
        parity = {"what": "step-0 eval-mode logits and CE loss of slide 0 vs the CPU oracle (same initial parameters)",
                  "loss_gpu": loss_gpu, "loss_oracle": loss_o, "loss_rel_err": abs(loss_gpu - loss_o) / abs(loss_o),
                  "logits_rel_err": err_l, "tol": tol}
        assert parity["loss_rel_err"] <= tol and err_l <= tol, "bench parity probe failed: %s" % parity
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # samples from the last warm-up steps through the timed and e2e regions (all under load)
    for _ in range(max(args.warmup, 3)):
        wl.step(wl.x_dev, wl.y_dev)
    l0 = sml_b200._lib.launch_count()
    wl._step(wl.x_dev, wl.y_dev)                             # kernels of ONE step, counted on an eager step (a graph replay is
    launches_per_step = sml_b200._lib.launch_count() - l0    # one host call; its nodes are these same kernels)
    wl.enable_graph()
    for _ in range(2):
        wl.step(wl.x_dev, wl.y_dev)
    ms = timed(lambda: wl.step(wl.x_dev, wl.y_dev), args.steps)
    launches = launches_per_step * args.steps
    launch_note = getattr(wl, "graph_note", "eager")

    # ---- end to end through the repo's loader (sml_b200.data.bags): the rank's bags sit in a page-locked host shard in the
    #      feed format of the mode (bf16-packed in "bf16" mode, fp32 otherwise); BagLoader copies every step's batch H2D on a side
    #      stream (prefetch 2, inside the timed region) and the loss is read back (D2H) every step.
    from sml_b200.data import BagLoader, BagShard
    if wl.x_dev.dtype == torch.bfloat16:
        shard = BagShard.from_tensors(wl.x_dev, labels=wl.y_dev)
        loader = BagLoader(shard, B, dev, shuffle=False, prefetch=2)
        h2d = loader.bytes_per_batch

        def batches(n):
            for xb_, _, _, _, yb_ in loader.stream(n):
                yield xb_, yb_
    else:                                                   # fp32-grade modes keep the reference's fp32 feed
        x_host = wl.x_dev.cpu().pin_memory(); y_host = wl.y_dev.cpu().pin_memory()
        h2d = int(x_host.numel() * 4 + y_host.numel() * 8)
        copy_stream = torch.cuda.Stream(device=dev)
        xb = [torch.empty_like(wl.x_dev) for _ in range(2)]
        yb = [torch.empty_like(wl.y_dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def batches(n):
            def issue(slot):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[slot])
                    xb[slot].copy_(x_host, non_blocking=True); yb[slot].copy_(y_host, non_blocking=True)
                    ready[slot].record(copy_stream)
            for sl in range(2):
                freed[sl].record(torch.cuda.current_stream())
            issue(0)
            for i in range(n):
                slot = i & 1
                if i + 1 < n:
                    issue(slot ^ 1)
                torch.cuda.current_stream().wait_event(ready[slot])
                yield xb[slot], yb[slot]
                freed[slot].record(torch.cuda.current_stream())

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def e2e_run(n):
        # every step's loss is copied device -> host (pinned, asynchronous) and read by the host one step later, the way a
        # training loop logs it: the D2H read of step i overlaps step i + 1 instead of draining the GPU every step; the
        # last loss is read before the timed region ends
        i = 0
        for xb_, yb_ in batches(n):
            loss = wl.step(xb_, yb_).detach()
            loss_host[i & 1].copy_(loss.float().reshape(()), non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                losses.append(float(loss_host[(i - 1) & 1]))
            i += 1
        loss_ev[(i - 1) & 1].synchronize()
        losses.append(float(loss_host[(i - 1) & 1]))
    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_t)
    sampler.stop_flag = True

    roof = instrumented_roofline(wl, ms / args.steps, pk, pk_src, args.prof_out if rank == 0 else None)

    # ---- secondary workloads (same N ranks): the north_star model (configs[2]), the loss + gather + sync-BN exchange at
    #      global batch 64 (configs[3]) and the fp32-grade mode of the headline (configs[1] "fp32 and bf16")
    secondary = []
    names = [n for n in args.secondary.split(",") if n] if args.workload == "transmil" else []
    import gc
    x_keep = (wl.x_dev, wl.y_dev)
    del wl
    gc.collect()                       # the graphed step holds the workload in a reference cycle: without the collection the
    torch.cuda.empty_cache()           # graph's private pool of every earlier workload stays allocated (OOM during a later capture)
    for nm in names:
        wname, prec, b2 = {"transmil_bf16x3": ("transmil", "bf16x3", B), "pathomic": ("pathomic", args.precision, B),
                           "pathomic1d": ("pathomic1d", args.precision, B),
                           "batchloss64": ("batchloss64", "fp32", max(64 // world, 1)),
                           "stress100k": ("stress100k", args.precision, 1)}[nm]
        w2 = Workload(wname, prec, b2, N, dev, rank, world, local, graph=args.graph)
        if wname not in ("batchloss64", "stress100k") and prec == args.precision:
            w2.x_dev, w2.y_dev = x_keep                       # same bags (and feed format) as the headline workload
        for _ in range(3):
            w2.step(w2.x_dev, w2.y_dev)
        w2.enable_graph()
        w2.step(w2.x_dev, w2.y_dev)
        k2 = args.secondary_steps * (20 if wname == "batchloss64" else (4 if wname == "stress100k" else 1))
        ms2 = timed(lambda: w2.step(w2.x_dev, w2.y_dev), k2)
        ent = {"workload": ("BatchLoss + PackedGather(all ranks) + BilinearFusion with sync-BN (train) fwd+bwd+Adam, global batch %d"
                            % (b2 * world)) if wname == "batchloss64" else
                           ("TransMIL inference (eval, no grad) on 100 489-patch x 1024-d bags, 512 landmarks, one bag per GPU (replicas)"
                            if wname == "stress100k" else workload_name(wname, N)),
               "baseline_config": {"transmil_bf16x3": "configs[1] (fp32-grade)", "pathomic": "configs[2] (YAML default: attn_dim 2 + BatchLoss)",
                                   "pathomic1d": "configs[2] (as written: DeformableAttention1D)", "batchloss64": "configs[3]",
                                   "stress100k": "configs[4]"}[nm],
               "precision": prec, "slides_per_step_per_gpu": b2, "value": b2 * world * k2 / (ms2 * 1e-3), "unit": UNIT,
               "ms_per_step": ms2 / k2, "steps": k2, "warmup": 3, "launch": getattr(w2, "graph_note", "eager")}
        if wname not in ("batchloss64", "stress100k"):
            r2 = instrumented_roofline(w2, ms2 / k2, pk, pk_src, None)
            if r2:
                ent["roofline"] = {k: r2[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "share_of_step") if k in r2}
        secondary.append(ent)
        del w2
        gc.collect()
        torch.cuda.empty_cache()

    if rank == 0:
        total = B * world * args.steps
        line = {"metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "f32 (bf16x3 split)", "fp32": "f32"}[args.precision],
                "data": "synthetic", "config": workload_config(args, world),
                "gpu_launches": int(launches), "launch": launch_note,
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": 4},
                "feed": ("bf16-packed bags in a page-locked shard, sml_b200.data.BagLoader (prefetch 2)" if args.precision == "bf16"
                         else "fp32 bags in pinned host memory, double-buffered side-stream copies"),
                "clocks": sampler.summary(), "roofline": roof, "parity": parity, "secondary": secondary}
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


def instrumented_roofline(wl, ms_step, pk, pk_src, prof_out):
    """per-kernel roofline: one instrumented step (CUDA events around every GEMM / fused-attention launch on the launch stream)"""
    from sml_b200 import ops
    ops.PROFILE = []
    wl._step(wl.x_dev, wl.y_dev)                             # eager pass (a graph replay would bypass the Python-side event hooks)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg, shapes = {}, {}
    for eng, flops, e0, e1, tag in prof:
        dt = e0.elapsed_time(e1) * 1e-3
        a = agg.setdefault(eng, [0.0, 0.0, 0])
        a[0] += flops; a[1] += dt; a[2] += 1
        sh = shapes.setdefault((eng,) + tag, [0.0, 0.0, 0])
        sh[0] += flops; sh[1] += dt; sh[2] += 1
    if prof_out:
        with open(prof_out, "w") as f:
            f.write("# eng M N K batch A_MNmajor B_MNmajor split relu|2softmax C Cb_hi Cb_lo beta : launches total_us avg_us TFLOP/s\n")
            for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
                f.write("%s : %d %.1f %.1f %.1f\n" % (" ".join(str(x) for x in k), v[2], v[1] * 1e6, v[1] * 1e6 / v[2], v[0] / v[1] / 1e12))
    # per-shape roofline of the heaviest tcgen05 GEMM shapes: bound by their outputs (HBM) when K = 64, by the tensor pipe
    # otherwise.  `traffic` = dram read + write per launch from the round's `ncu --set full` capture (profiles/rNN_ncu_traffic.json)
    ncu_traffic, traffic_src = load_ncu_traffic()
    by_shape = []
    tr_sum, tr_n = 0.0, 0
    for key, v in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
        if key[0] != "tc" or len(key) < 13:
            continue
        M_, N_, K_, nb_, _, _, split_, relu_, c_, hi_, lo_, _ = key[1:]
        tr = ncu_traffic.get((M_, N_, K_, nb_, relu_ & 2))
        if tr is not None:
            tr_sum += tr * v[2]; tr_n += v[2]
        if len(by_shape) >= 5:
            continue
        t_avg = v[1] / v[2]
        bytes_ = nb_ * (2.0 * (1 + split_) * (M_ * K_ + K_ * N_) + M_ * N_ * (4.0 * c_ + 2.0 * hi_ + 2.0 * lo_))
        hbm_frac = bytes_ / t_avg / 1e9 / pk["hbm_gbs"]
        tc_frac = v[0] / v[1] / 1e12 / pk["bf16_tflops_sustained"]
        by_shape.append({"shape": "%dx%dx%d x%d%s%s" % (M_, N_, K_, nb_, " bf16x3" if split_ else "", " +softmax" if relu_ & 2 else ""),
                         "launches": v[2], "avg_us": t_avg * 1e6,
                         "bound": "latency (dependent chain, L2-resident operands)" if (K_ == M_ == N_ == 256) else
                                  ("hbm" if hbm_frac > tc_frac else "tensor"),
                         "achieved": (bytes_ / t_avg / 1e9) if hbm_frac > tc_frac else v[0] / v[1] / 1e12,
                         "unit": "GB/s" if hbm_frac > tc_frac else "TFLOP/s", "frac": max(hbm_frac, tc_frac), "traffic": tr})
    roof = None
    peak = pk["bf16_tflops_sustained"]
    if wl.deform and "deform_bwd" in agg:
        f, t, n = agg["deform_bwd"]
        ach = f / t / 1e12
        roof = {"kernel": "deform_attn_bwd kernel (fused bias-MLP + attention backward, %d launch)" % n, "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": pk_src + " (sustained, tcgen05 dense bf16)", "share_of_step": t / (ms_step * 1e-3),
                "forward_kernel": {"achieved": agg["deform_fwd"][0] / agg["deform_fwd"][1] / 1e12,
                                   "share_of_step": agg["deform_fwd"][1] / (ms_step * 1e-3)},
                "note": "algorithmic FLOPs: 2*(2240 (2-D) | 2176 (1-D) + 256) per (query, key) pair for the backward; the in-kernel "
                        "forward recomputation and the 3-pass split are not counted.  1-D bf16 mode: the bias MLP is an exact "
                        "piecewise-linear table (no per-pair MLP), counted as 10*64 + 24 FLOP per pair backward, 4*64 + 8 forward"}
    elif "tc" in agg and agg["tc"][1] > 0:
        f, t, n = agg["tc"]
        ach = f / t / 1e12
        roof = {"kernel": "gemm_tc_kernel (tcgen05, all %d launches of one step)" % n, "bound": "tensor", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": (tr_sum / tr_n) if tr_n else None, "traffic_note": ("mean dram read + write bytes per launch over the %d of %d "
                "launches whose shape is in the ncu capture %s" % (tr_n, n, traffic_src)) if tr_n else None,
                "peak_source": pk_src + " (sustained)", "share_of_step": t / (ms_step * 1e-3),
                "note": "bf16x3 launches count 1x algorithmic FLOPs (3 MMA passes executed)", "by_shape": by_shape}
        # the flash-fused landmark-by-token kernels (csrc/nystrom_a3.cu: warp-level mma.sync, not part of the GEMM engine's sum)
        fused = {}
        for eng in ("a3_fwd", "a3_bwd"):
            if eng in agg and agg[eng][1] > 0:
                f2, t2, n2 = agg[eng]
                fused[eng] = {"launches": n2, "avg_us": t2 / n2 * 1e6, "achieved": f2 / t2 / 1e12, "unit": "TFLOP/s",
                              "frac": f2 / t2 / 1e12 / peak, "share_of_step": t2 / (ms_step * 1e-3),
                              "flop_per_pair": 256 if eng == "a3_fwd" else 640}
        if fused:
            roof["fused_attention"] = fused
    elif "simt" in agg:
        f, t, n = agg["simt"]
        roof = {"kernel": "gemm_f32_kernel (SIMT fp32, %d launches)" % n, "bound": "tensor", "achieved": f / t / 1e12,
                "peak": peak, "unit": "TFLOP/s", "frac": f / t / 1e12 / peak, "traffic": None, "peak_source": pk_src}
    return roof


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
