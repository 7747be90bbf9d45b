"""GPU comparator (SURVEY.md §8d "the real bar"): the reference's algorithm as plain PyTorch eager ops on the SAME B200.

The reference modules cannot travel to the GPU box (/root/reference does not exist there) and ship no Blackwell kernel of
their own: what "the reference on this GPU" executes is ATen / cuDNN 9 kernels called op by op.  The oracle restatement
(pinned to the reference modules at max |diff| = 0, oracle/validate_against_reference.py) issues exactly those ops, so it
is timed here on cuda:0 as the comparator -- in the precisions a user could pick: fp32 with TF32 (the reference's default
on Ampere+ for convs), and bf16 autocast.  Test infrastructure: lives under tests/, not imported by the product.

    python tests/bench_torch_eager_gpu.py [--train]

Prints ms per guided denoising step (UNet batch 64 = 32 objects x CFG) and, with --train, per fwd+bwd at batch 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import denoiser as D, weights as Wt


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cfg = D.UNET_FULL
    sd = {k: v.cuda() for k, v in Wt.synth_state_dict(D.unet_param_shapes(cfg), 111).items()}
    g = torch.Generator().manual_seed(0)
    B = 64
    x = torch.randn(B, 3, 16, 16, 16, generator=g).cuda()
    t = torch.full((B,), 500).cuda()
    ctx = torch.randn(B, 1, 1280, generator=g).cuda()
    torch.backends.cudnn.benchmark = True
    rows = []
    with torch.device("cuda"), torch.no_grad():
        for name, tf32, ac in (("fp32 (TF32 convs+matmuls)", True, False), ("fp32 (no TF32)", False, False), ("bf16 autocast", True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                ms = timed(lambda: D.unet_forward(sd, cfg, x, t, ctx), 3 if tf32 else 1)
            rows.append((name, ms))
            print(f"torch eager UNet forward, batch 64, {name}: {ms:.1f} ms -> {1000 / ms:.2f} guided steps/s "
                  f"({64 * 557.6 / ms:.0f} TFLOP/s)", flush=True)
    if "--concat" in sys.argv:       # the concat-conditioning variant (AttentionBlock, in_channels 4, isotropic resampling)
        ccfg = D.UNET_CONCAT_FULL
        csd = {k: v.cuda() for k, v in Wt.synth_state_dict(D.unet_param_shapes(ccfg), 111).items()}
        cc = torch.randn(B, 1, 16, 16, 16, generator=g).cuda()
        with torch.device("cuda"), torch.no_grad():
            for name, tf32, ac in (("fp32 (TF32 convs+matmuls)", True, False), ("bf16 autocast", True, True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    ms = timed(lambda: D.unet_forward(csd, ccfg, x, t, c_concat=cc), 3)
                print(f"torch eager CONCAT UNet forward, batch 64, {name}: {ms:.1f} ms -> {1000 / ms:.2f} guided steps/s", flush=True)
    if "--train" in sys.argv:
        B = 32
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x, t, ctx = x[:B], t[:B], ctx[:B]
        noise = torch.randn_like(x)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True

        def step(ac):
            with torch.device("cuda"), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                eps = D.unet_forward(sdg, cfg, x, t, ctx)
            loss = torch.nn.functional.mse_loss(eps.float(), noise)
            torch.autograd.grad(loss, list(sdg.values()), allow_unused=True)
        for name, ac in (("fp32 (TF32)", False), ("bf16 autocast", True)):
            ms = timed(lambda: step(ac), 2)
            print(f"torch eager UNet fwd+bwd (no activation checkpointing, no optimizer), batch 32, {name}: {ms:.1f} ms "
                  f"({3 * 32 * 557.6 / ms:.0f} TFLOP/s), peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


if __name__ == "__main__":
    main()
