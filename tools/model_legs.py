"""Model-level legs of bench.py (configs 1, 3, 4, 5 of BASELINE.json) on the UNMODIFIED reference ITS model staged under
baseline/_ref (baseline/fetch_its.sh) with focalnet_b200.patch_ss2d applied.  Bench infrastructure, not product.

  train_leg : config 3 (1 GPU) / config 5 (N GPUs, data parallel): 32 synthetic 256x256 crops per GPU, one full step of
              ITS/train.py:57-91 per iteration — forward, 6-term loss, backward, gradient all-reduce over NCCL (N > 1,
              overlapped with the backward's tail), global-norm clip 0.001, Adam.  img/s = N * 32 * steps / max-over-ranks
              device time.  At N = 1 the same step is also timed on the reference's own kernels (oflex CUDA rebuilt for
              sm_100a + the shipped Triton CrossScan / CrossMerge + torch clip / Adam).
  infer_leg : config 4: g4 model, T full-resolution 620x460 images per GPU, ITS/eval.py:33-41 (reflect pad to 640x480,
              no_grad), images sharded over the ranks; PSNR against the reference kernels at N = 1.
  cpu_forward_leg : config 1: batch-1 256x256 forward on the host cores through selective_scan_ref.
"""
from __future__ import annotations

import os
import time

import torch

CROPS_PER_GPU = 32
TILES_PER_GPU = 8


def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def _timed(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    a, b = _events(2)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    barrier()
    return a.elapsed_time(b) / steps


def _bind_reference(H, model):
    try:
        H.bind_reference_cuda(model, triton_cross=True)
        with torch.no_grad():
            model(torch.rand(1, 3, 32, 32, device="cuda"))
        return "reference oflex CUDA (sm_100a rebuild) + Triton CrossScan/CrossMerge (forward type v4 as shipped)"
    except Exception:
        H.bind_reference_cuda(model, triton_cross=False)
        return "reference oflex CUDA (sm_100a rebuild) + torch CrossScan/CrossMerge (Triton JIT unavailable)"


def train_leg(dev, rank, world, steps, warmup, barrier, max_over_ranks, with_reference=True):
    from baseline import its_harness as H
    from focalnet_b200 import patch_ss2d
    from focalnet_b200.dp import FlatBucket, FusedClipAdam, dp_train_step, expected_allreduce_bytes
    if not H.available():
        return {"unavailable": "reference model files not staged (baseline/fetch_its.sh)"}
    model = H.build_model("g2", "cuda")      # seed 1234 on every rank: identical replicas
    n_patched = patch_ss2d(model)
    model.train()
    bucket = FlatBucket(model.parameters(), segments=3)
    opt = FusedClipAdam(bucket, lr=1e-4, max_norm=0.001)
    x, J = H.synthetic_pair(CROPS_PER_GPU, 256, 256, dev, seed=100 + rank)
    torch.manual_seed(1000 + rank)           # DropPath masks differ per rank, like independent data-loader workers
    exposed = []

    def step():
        bucket.begin_step()
        loss = H.its_loss(model(x), J)
        loss.backward()
        a, b = _events(2)
        a.record()
        bucket.finish_reduce()
        b.record()
        exposed.append((a, b))
        opt.step()
        return loss.detach()

    for _ in range(warmup):
        step()
    barrier()
    exposed.clear()
    torch.cuda.reset_peak_memory_stats(dev)
    a, b = _events(2)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b) / steps, dev)
    out = {"img_s": world * CROPS_PER_GPU / (ms * 1e-3), "ms_per_step": ms, "crops_per_gpu": CROPS_PER_GPU, "image": "256x256",
           "steps": steps, "model": "MIMOUNet g2 (unchanged ITS/models, 2,541,673 params)", "ss2d_modules_patched": n_patched,
           "loss": float(loss), "grad_norm_before_clip": float(opt.last_norm), "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
           "step": "fwd + 6-term loss + bwd + allreduce + clip 0.001 + Adam (ITS/train.py:57-91)", "dtype": "f32 (TF32 convs as torch default)"}
    if world > 1:
        out["exposed_comm_ms"] = max_over_ranks(sum(p.elapsed_time(q) for p, q in exposed) / len(exposed), dev)
        bucket.zero_grad()
        out["allreduce_ms"] = max_over_ranks(_timed(bucket.allreduce_whole, 20, 5, barrier), dev)
        out["allreduce_bytes"] = expected_allreduce_bytes(bucket)
        out["allreduce_segments"] = len(bucket.seg_bounds)
        bucket.zero_grad()
    # end to end through the public API: every step copies its crops + labels from pinned host memory and reads the loss back
    hx, hJ = x.cpu().pin_memory(), J.cpu().pin_memory()

    def e2e_step():
        xd, Jd = hx.to(dev, non_blocking=True), hJ.to(dev, non_blocking=True)
        bucket.begin_step()
        loss = H.its_loss(model(xd), Jd)
        loss.backward()
        bucket.finish_reduce()
        opt.step()
        return loss.item()  # D2H of the step's result

    n_e2e = max(3, min(steps, 8))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e, dev)
    out["e2e"] = {"img_s": world * CROPS_PER_GPU / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                  "h2d_bytes_per_step": 2 * hx.numel() * 4, "d2h_bytes_per_step": 4}
    del model, bucket, opt
    torch.cuda.empty_cache()
    if with_reference and world == 1 and rank == 0:
        try:
            ref = H.build_model("g2", "cuda")
            how = _bind_reference(H, ref)
            ref.train()
            ropt = H.make_optimizer(ref, lr=1e-4)
            torch.manual_seed(1000 + rank)
            rms = _timed(lambda: H.train_step(ref, ropt, x, J), max(5, min(steps, 10)), 3, barrier)
            out["reference_kernels"] = {"img_s": CROPS_PER_GPU / (rms * 1e-3), "ms_per_step": rms, "what": how,
                                        "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9}
            out["speedup_vs_reference_kernels"] = rms / ms
            del ref, ropt
        except Exception as exc:  # comparison arm only
            out["reference_kernels"] = {"unavailable": repr(exc)[:200]}
        torch.cuda.empty_cache()
    return out


def infer_leg(dev, rank, world, steps, warmup, barrier, max_over_ranks, with_reference=True):
    from baseline import its_harness as H
    from focalnet_b200 import patch_ss2d, unpatch_ss2d
    if not H.available():
        return {"unavailable": "reference model files not staged (baseline/fetch_its.sh)"}
    model = H.build_model("g4", "cuda")
    model.eval()
    x, J = H.synthetic_pair(TILES_PER_GPU, 460, 620, dev, seed=200 + rank)
    patch_ss2d(model)
    with torch.no_grad():
        ms = max_over_ranks(_timed(lambda: H.eval_forward(model, x), steps, warmup, barrier), dev)
        y = H.eval_forward(model, x)
    out = {"img_s": world * TILES_PER_GPU / (ms * 1e-3), "ms_per_batch": ms, "images_per_gpu": TILES_PER_GPU,
           "image": "620x460 reflect-padded to 640x480 (ITS/eval.py:33-37)", "model": "MIMOUNet g4 (results_1mlp_g4, patch_size_global=4)",
           "scan_L": [19200, 4800, 1200], "psnr_dB": H.psnr(y, J), "ssim": H.eval_metrics(y, J)[1]}
    # one image per call (ITS/eval.py:19 evaluates with batch_size=1): eager launches vs ONE CUDA-graph replay
    if world == 1 and rank == 0:
        try:
            from focalnet_b200 import GraphedForward
            x1 = x[:1].contiguous()
            with torch.no_grad():
                eager_ms = _timed(lambda: H.eval_forward(model, x1), max(5, steps), 3, barrier)
                gf = GraphedForward(lambda t: H.eval_forward(model, t), x1)
                graph_ms = _timed(lambda: gf(x1), max(5, steps), 3, barrier)
                same = float((gf(x1) - H.eval_forward(model, x1)).abs().max())
            out["single_image"] = {"eager_ms": eager_ms, "cuda_graph_ms": graph_ms, "img_s_cuda_graph": 1e3 / graph_ms,
                                   "max_abs_diff_graph_vs_eager": same}
            del gf
        except Exception as exc:
            out["single_image"] = {"error": repr(exc)[:200]}
    if with_reference and world == 1 and rank == 0:
        try:
            unpatch_ss2d(model)
            how = _bind_reference(H, model)
            with torch.no_grad():
                rms = _timed(lambda: H.eval_forward(model, x), max(3, min(steps, 10)), 2, barrier)
                y_ref = H.eval_forward(model, x)
            out["reference_kernels"] = {"img_s": TILES_PER_GPU / (rms * 1e-3), "ms_per_batch": rms, "what": how, "psnr_dB": H.psnr(y_ref, J),
                                        "ssim": H.eval_metrics(y_ref, J)[1]}
            out["psnr_delta_dB"] = abs(out["psnr_dB"] - H.psnr(y_ref, J))
            if "single_image" in out and "eager_ms" in out["single_image"]:
                with torch.no_grad():
                    out["single_image"]["reference_kernels_eager_ms"] = _timed(lambda: H.eval_forward(model, x[:1].contiguous()), max(5, steps), 3, barrier)
            out["speedup_vs_reference_kernels"] = rms / ms
        except Exception as exc:
            out["reference_kernels"] = {"unavailable": repr(exc)[:200]}
    del model
    torch.cuda.empty_cache()
    return out


def cpu_forward_leg(size=256):
    """Config 1: MIMOUNet (g2) forward, batch 1, one synthetic size x size hazy image, on the host cores with torch
    CrossScan / CrossMerge and selective_scan_ref (the oracle's torch port)."""
    from baseline import its_harness as H
    if not H.available():
        return {"unavailable": "reference model files not staged (baseline/fetch_its.sh)"}
    torch.set_num_threads(os.cpu_count() or 1)
    model = H.build_model("g2", "cpu")
    H.bind_cpu_reference(model)
    model.eval()
    x, J = H.synthetic_pair(1, size, size, "cpu", seed=300)
    t0 = time.perf_counter()
    with torch.no_grad():
        y = model(x)[2]
    dt = time.perf_counter() - t0
    return {"img_s": 1.0 / dt, "seconds": dt, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "image": f"{size}x{size}",
            "kind": "port", "what": "unchanged MIMOUNet g2, torch CrossScan/CrossMerge + selective_scan_ref (oracle torch port), fp32",
            "psnr_dB": H.psnr(y, J)}
