#!/usr/bin/env python
"""bench.py — 512x1024 training crops/sec of the MDIL-SS hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload step1|step2]

A "step" is one pass of the hot path over one batch of synthetic Cityscapes-shaped input:
  step1 (default, BASELINE configs[1]): Step-1 single-domain train, 20 classes, batch 6 per GPU, 512x1024:
        forward (train BN, dropout) -> CrossEntropy2d -> backward -> [one gradient all-reduce] -> Adam.
  step2 (configs[2] shape per GPU): 2-domain adapters + KD: student fwd x2, teacher fwd, CE + 0.1*KD, backward, Adam.
Per-GPU work is fixed as N grows (weak scaling); value = crops all ranks processed / max-over-ranks device time.
`--impl reference` times the reference algorithm's CPU implementation (the oracle port: the reference is pure
Python/PyTorch and /root/reference does not exist on the GPU box) on a bounded sample, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

H, W, NCLS, BATCH_PER_GPU = 512, 1024, 20, 6
METRIC = "train_crops_per_sec_512x1024"
UNIT = "crops/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="step1", choices=["step1", "step2", "step3", "multitask"])
    ap.add_argument("--full-res", action="store_true", help="1024x2048 crops (BASELINE config 5) instead of 512x1024")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="crops per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch-eager (cuDNN) baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the captured step")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: after warm-up run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    return ap.parse_args()


def workload_name(args):
    if args.workload == "step1":
        return (f"Step-1 CS single-domain train (fwd+CE2d+bwd+allreduce+Adam), 20 cls, batch {args.batch}/GPU, "
                f"{H}x{W} synthetic")
    if args.workload == "step2":
        return (f"Step-2 CS->BDD train (student fwd x2 + teacher fwd, CE2d + 0.1*KD, bwd, allreduce, Adam), 20/20 cls, "
                f"batch {args.batch}/GPU, {H}x{W} synthetic")
    if args.workload == "multitask":
        return (f"Multi-task joint CS+BDD+IDD over the RAP network (one visit per dataset and step: fwd, CE2d, bwd, "
                f"allreduce, Adam), 20/20/27 cls, batch {args.batch}/GPU and dataset, {H}x{W} synthetic")
    return (f"Step-3 CS|BDD->IDD train (CE step, then student fwd x2 + teacher fwd x2 + 0.1*(KD+KD) step: two "
            f"all-reduce + Adam per iteration), 20/20/27 cls, batch {args.batch}/GPU, {H}x{W} synthetic")


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active") and not v.lower().startswith("not"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def oracle_step_factory(workload, n):
    import torch
    from oracle import erfnet_rap_oracle as oracle
    g = torch.Generator().manual_seed(1234)
    images = torch.rand(n, 3, H, W, generator=g)
    labels = torch.randint(0, 27 if workload in ("step3", "multitask") else NCLS, (n, 1, H // 32, W // 32),
                           generator=g).repeat_interleave(32, 2).repeat_interleave(32, 3)
    weight = torch.tensor({"step1": oracle.WEIGHT_CITY, "step2": oracle.WEIGHT_BDD, "step3": oracle.WEIGHT_IDD,
                           "multitask": oracle.WEIGHT_IDD}[workload])
    if workload == "step1":
        sd = oracle.init_state_dict([NCLS], 1, seed=0)
        state = [dict() for _ in oracle.param_names(sd)]

        def step():
            torch.manual_seed(7)
            noise = oracle.make_dropout_noise(n, True)
            loss, _, _ = oracle.step1_iteration(sd, images, labels, weight, 0, noise, state)
            return float(loss)
    elif workload == "multitask":
        sd = oracle.init_state_dict([NCLS, NCLS, 27], 3, seed=0)
        state = {}
        ws = [torch.tensor(w) for w in (oracle.WEIGHT_CITY, oracle.WEIGHT_BDD, oracle.WEIGHT_IDD)]
        lab20 = labels.clamp(max=NCLS - 1)

        def step():
            torch.manual_seed(7)
            noises = [oracle.make_dropout_noise(n, True) for _ in range(3)]
            losses = oracle.multitask_iteration(sd, [(images, lab20), (images, lab20), (images, labels)], ws, noises, state)
            return float(losses[-1])
    elif workload == "step3":
        sd_old = oracle.init_state_dict([NCLS, NCLS], 2, seed=0)
        sd = oracle.init_state_dict([NCLS, NCLS, 27], 3, seed=1)
        state = {}

        def step():
            torch.manual_seed(7)
            noises = [oracle.make_dropout_noise(n, True) for _ in range(5)]
            ce, kd, _, _, _ = oracle.step3_iteration(sd, sd_old, images, labels, weight, 2, 0.1, noises, state)
            return float(ce) + float(kd)
    else:
        sd_old = oracle.init_state_dict([NCLS], 1, seed=0)
        sd = oracle.init_state_dict([NCLS, NCLS], 2, seed=1)
        names = oracle.trainable_names_incremental(sd, 1)
        state = [dict() for _ in names]

        def step():
            torch.manual_seed(7)
            n1 = oracle.make_dropout_noise(n, True)
            n2 = oracle.make_dropout_noise(n, True)
            ce, kd, _, _ = oracle.step2_iteration(sd, sd_old, images, labels, weight, 1, 0.1, n1, n2, state)
            return float(ce) + 0.1 * float(kd)
    return step


def time_cpu(workload, n, steps, warmup):
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread (set before torch spins up its pool)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ["MKL_NUM_THREADS"] = str(cores)
    import torch
    torch.set_num_threads(cores)
    step = oracle_step_factory(workload, n)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps * 1e3, cores


def time_gpu_eager(n, steps, warmup, allow_tf32, dev):
    """Like-for-like GPU baseline: the reference ALGORITHM as the reference runs it on a
    GPU -- PyTorch eager, ATen/cuDNN convolutions and batch norm, NCHW fp32, cudnn.benchmark off (the drivers never set
    it), F.nll_loss(F.log_softmax) as CrossEntropyLoss2d (train_new_task_step2.py:84-92), torch.optim.Adam as the drivers
    configure it (:237-239) -- on the same B200, same synthetic step-1 batch.  Functional restatement = the oracle port
    moved to the device (the unmodified reference files do not exist on the GPU box)."""
    import torch
    import torch.nn.functional as F
    from oracle import erfnet_rap_oracle as oracle
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = bool(allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    try:
        g = torch.Generator().manual_seed(1234)
        images = torch.rand(n, 3, H, W, generator=g).to(dev)
        labels = (torch.randint(0, NCLS, (n, H // 32, W // 32), generator=g)
                  .repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous().to(dev))
        weight = torch.tensor(oracle.WEIGHT_CITY, device=dev)
        sd = {k: v.to(dev) for k, v in oracle.init_state_dict([NCLS], 1, seed=0).items()}
        names = oracle.param_names(sd)
        for k in names:
            sd[k].requires_grad_(True)
        opt = torch.optim.Adam([sd[k] for k in names], 5e-4, (0.9, 0.999), eps=1e-08, weight_decay=1e-4)

        def step():
            # Dropout2d noise drawn on the device, as F.dropout2d does inside the reference's blocks
            noise = [torch.empty(n, ch, 1, 1, device=dev).bernoulli_(1 - p).div_(1 - p) if kind == "rap" and p != 0 else None
                     for kind, ch, p, _ in oracle.ENCODER_LAYERS]
            logits = oracle.net_forward(sd, images, 0, True, noise)
            opt.zero_grad()
            loss = F.nll_loss(F.log_softmax(logits, dim=1), labels, weight)
            loss.backward()
            opt.step()
            return loss

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del sd, opt
        torch.cuda.empty_cache()
        return n * 1e3 / ms, ms
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 1
    steps = max(1, min(args.steps, 4))
    warmup = 1
    value, ms, cores = time_cpu(args.workload, n, steps, warmup)
    sample = f"{steps} timed + {warmup} warm-up iterations of the same step at batch {n} (per-crop CPU cost is flat in N)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "arm": "reference algorithm on host CPU (oracle port, PyTorch ATen/oneDNN, fp32)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from mdil_ss_b200 import _lib
    from mdil_ss_b200.erfnet_RA_parallel import Net
    from mdil_ss_b200.parallel import broadcast_module
    from mdil_ss_b200.train_step import Step1Trainer, Step2Trainer, class_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    if lib.mdil_device_supported(local_rank) != 1:
        raise SystemExit("bench.py: libmdil_b200.so targets sm_100a (B200) only")

    n = args.batch
    torch.manual_seed(0)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        if args.workload == "step1":
            model = Net([NCLS], 1, 0).to(dev)
        elif args.workload == "step2":
            model_old = Net([NCLS], 1, 0).to(dev)
            model = Net([NCLS, NCLS], 2, 1).to(dev)
        elif args.workload == "multitask":
            model = Net([NCLS, NCLS, 27], 3, 0).to(dev)
        else:
            model_old = Net([NCLS, NCLS], 2, 1).to(dev)
            model = Net([NCLS, NCLS, 27], 3, 2).to(dev)
    broadcast_module(model)
    ncls_labels = 27 if args.workload == "multitask" else NCLS
    if args.workload == "step1":
        trainer = Step1Trainer(model, class_weights("cityscapes", dev))
    elif args.workload == "step2":
        broadcast_module(model_old)
        trainer = Step2Trainer(model, model_old, class_weights("BDD", dev), 1, 0.1)
    elif args.workload == "multitask":
        from mdil_ss_b200.train_step import MultiTaskTrainer
        trainer = MultiTaskTrainer(model, [class_weights(d_, dev) for d_ in ("cityscapes", "BDD", "IDD")])
    else:
        from mdil_ss_b200.train_step import Step3Trainer
        broadcast_module(model_old)
        trainer = Step3Trainer(model, model_old, class_weights("IDD", dev), 2, 0.1)
        ncls_labels = 27

    g = torch.Generator().manual_seed(1234 + rank)
    images_h = torch.rand(n, 3, H, W, generator=g).pin_memory()
    labels_h = (torch.randint(0, ncls_labels, (n, 1, H // 32, W // 32), generator=g)
                .repeat_interleave(32, 2).repeat_interleave(32, 3).contiguous().pin_memory())
    images = images_h.to(dev, non_blocking=True)
    labels = labels_h.to(dev, non_blocking=True)

    crops_per_step = n
    if args.workload == "multitask":      # three dataset visits per step (labels of the 20-class sets clipped to their range)
        crops_per_step = 3 * n
        _orig_step = trainer.step
        trainer.step = lambda x, y: _orig_step([(x, y.clamp(max=NCLS - 1)), (x, y.clamp(max=NCLS - 1)), (x, y)])[-1]

    def step_resident():
        return trainer.step(images, labels)

    from mdil_ss_b200.data import DevicePrefetcher
    prefetch = DevicePrefetcher(dev)
    prefetch.put(images_h, labels_h)

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

    for _ in range(max(3, args.warmup)):
        step_resident()
    if args.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- kernel launches of one step, and the per-kind CUDA-event timing of the fused pair kernels (roofline leg):
    # a few eager steps with the library's launch profiling on
    launches0 = _lib.launches()
    lib.mdil_profile_begin()
    prof_steps = 5
    barrier()
    for _ in range(prof_steps):
        step_resident()
    barrier()
    nk = 12
    tot = (ctypes.c_float * nk)()
    cnt = (ctypes.c_int * nk)()
    lib.mdil_profile_end(tot, cnt, nk)
    launches_per_step = (_lib.launches() - launches0) // prof_steps

    # ---- the step as the product runs it: captured once in a CUDA graph and replayed (one host call per step); falls
    # back to host-side launches where the trainer cannot be captured (step 3: per-parameter optimiser step counts)
    graphed = None
    if not args.no_graph and args.workload in ("step1", "step2"):
        try:
            from mdil_ss_b200.train_step import GraphedStep
            graphed = GraphedStep(trainer, images, labels)
        except Exception as exc:   # noqa: BLE001
            if rank == 0:
                print(f"bench.py: CUDA-graph capture unavailable ({exc!r}); launching from the host", file=sys.stderr)
            graphed = None
    graph_flags = torch.tensor([1.0 if graphed is not None else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(graph_flags, op=dist.ReduceOp.MIN)     # every rank or none (the collectives must line up)
    if float(graph_flags) < 1.0:
        graphed = None

    def run_step(x, y):
        return graphed.step(x, y) if graphed is not None else trainer.step(x, y)

    def step_value():
        return run_step(images, labels)

    # device -> host read of every step's result WITHOUT stalling the pipeline: the loss of step k is copied into a pinned
    # host word asynchronously and read (after its event) while step k+1 is already enqueued -- the drivers' `.item()`
    # only feeds a running average for the progress line (train_new_task_step2.py:305-312)
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"k": 0, "last": float("nan")}

    def step_e2e():
        # every step consumes a batch copied from pinned host memory; the copy of the NEXT batch is issued on the
        # prefetcher's side stream before this step's kernels, as a DataLoader-fed training loop would
        x, y = prefetch.get()
        prefetch.put(images_h, labels_h)
        out = run_step(x, y)
        loss = out[0] if isinstance(out, tuple) else out
        k = e2e_state["k"]
        loss_host[k & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ev[k & 1].record()
        if k > 0:                                   # read the previous step's loss: its copy finished long ago
            loss_ev[(k - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_host[(k - 1) & 1])
        e2e_state["k"] = k + 1
        return e2e_state["last"]

    for _ in range(2):
        step_value()
    # ---- timed region (device-resident inputs), clocks sampled during it
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_value, args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * crops_per_step * args.steps / (ms_total / 1e3)
    # tot[] / cnt[] cover prof_steps eager steps
    for i in range(nk):
        tot[i] = tot[i] * args.steps / prof_steps
        cnt[i] = cnt[i] * args.steps // prof_steps

    # ---- end-to-end: host buffers, H2D of the inputs and D2H of the loss inside the timed region
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * crops_per_step * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") \
            else (6650.0, "fallback (B200_PROFILING.md)")
        # dominant fused-block kernel: the kind with the largest total time
        names = [f"nb1d_pair<C={c}> {ph}" for c in (16, 64, 128) for ph in ("fwd pair1", "fwd pair2", "bwd pair2", "bwd pair1")]
        kind = max(range(nk), key=lambda i: tot[i])
        c = (16, 64, 128)[kind // 4]
        hh, ww = {16: (H // 2, W // 2), 64: (H // 4, W // 4), 128: (H // 8, W // 8)}[c]
        T = 4.0 * n * c * hh * ww
        # ALGORITHMIC bytes of one launch (SURVEY 8d / DESIGN.md 4), T = one tensor pass: forward pairs read their input
        # and write their output (2T); backward pair 2 = the "main pass" (R ds, c-mask, p; W dq: 4T); backward pair 1 =
        # the "input pass" (R dp, a-mask, dy, y; W dx: 5T).  The saved / gradient `mid` tensor a launch also writes is
        # not counted (it is an implementation choice: 180 GB of HBM make recomputation unnecessary).
        passes = (2.0, 2.0, 4.0, 5.0)[kind % 4]
        alg_bytes = passes * T
        avg_ms = tot[kind] / max(1, cnt[kind])
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        pair_share = sum(tot) / ms_total if ms_total > 0 else 0.0
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kind, from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(REPO, "profiles", "pair_traffic.json"))).get(names[kind])
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "kernel": names[kind], "avg_launch_ms": avg_ms, "launches_timed": int(cnt[kind]),
                    "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_tensor_passes": passes, "peak_source": peak_src,
                    "pair_kernels_share_of_step": pair_share,
                    "per_kind_ms_per_step": {names[i]: tot[i] / args.steps for i in range(nk) if cnt[i]}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args), "global_batch": world * n, "crop": f"{H}x{W}",
                           "parallelism": f"dp{world}", "l2": "per-step activations (>2 GB) exceed the 126 MB L2",
                           "collective": "one NCCL all-reduce of the flat fp32 gradient buffer per optimiser step (+ a 16-byte all-reduce of the CE accumulators: DataParallel's global loss normalisation)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": images_h.numel() * 4 + labels_h.numel() * 8, "d2h_bytes_per_step": 4,
                        "note": "every step: H2D of its batch from pinned memory (prefetched one step ahead on a side stream) and an asynchronous D2H of its loss, consumed by the host one step later"},
                "gpu_launches": int(launches), "cuda_graph": graphed is not None,
                "gpu_launches_note": "kernels of this library inside the timed region = launches of one step (counted on an eager step) x steps; with cuda_graph the step is replayed from one captured graph",
                "roofline": roofline}
        if not args.no_cpu_baseline and world == 1:   # CPU baseline: rank 0 at N=1 only
            cpu_steps = 2 if args.workload == "step1" else 1
            v, ms, cores = time_cpu(args.workload, 1, cpu_steps, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{cpu_steps} timed + 1 warm-up iterations of the same step at batch 1 on the host CPU "
                                              f"({ms:.0f} ms/iteration)"}
        if not args.no_gpu_baseline and world == 1 and args.workload == "step1" and not args.full_res:
            # like-for-like GPU baseline: the same step in PyTorch eager (ATen/cuDNN) on this B200, fp32 and TF32
            gb = {"kind": "reference algorithm (oracle port) on cuda: PyTorch eager, ATen/cuDNN NCHW, torch.optim.Adam; "
                          "cudnn.benchmark off as in the drivers", "unit": UNIT, "batch": n, "warmup": 10, "steps": 20}
            try:
                for tf32 in (False, True):
                    v, ms = time_gpu_eager(n, 20, 10, tf32, dev)
                    gb["allow_tf32=%s" % tf32] = {"value": v, "ms_per_step": ms}
                gb["value"] = gb["allow_tf32=False"]["value"]
                gb["allow_tf32"] = False
                gb["speedup_vs_fp32_eager"] = value / gb["allow_tf32=False"]["value"]
                gb["speedup_vs_tf32_eager"] = value / gb["allow_tf32=True"]["value"]
            except Exception as exc:   # the baseline must never take the bench line down
                gb["error"] = repr(exc)[:300]
            line["gpu_baseline"] = gb
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown: the captured graph holds NCCL kernels; destroying the process group under it was seen to hang
        # (2-GPU run, after the bench line had been printed).  Drop the graph, drain the device, and leave without the
        # collective teardown: every rank has passed the barrier, nothing is in flight.
        dist.barrier()
        graphed = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    global H, W, METRIC
    args = parse()
    if args.full_res:
        H, W = 1024, 2048
        METRIC = "train_crops_per_sec_1024x2048"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
