"""Generate the golden fixtures in tests/golden/*.npz from the REFERENCE's own modules.

Runs only in the build container (imports /root/reference through oracle/validate_against_reference.py's
shim).  Weights are the seeded synthetic tensors of oracle/weights.py written into the reference modules;
inputs are seeded CPU tensors.  The fixtures hold inputs (when small) and the reference's outputs, so that
on the GPU box — where /root/reference does not exist — both the oracle and the CUDA product can be checked
against what the reference itself computed.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import denoiser as D, graph as G, vqvae as V, weights as Wt  # noqa: E402
from oracle import validate_against_reference as R  # noqa: E402

SEED_UNET, SEED_VQ, SEED_GCN = 11, 13, 17


def unet_inputs(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    r = cfg["image_size"]
    x = torch.randn(B, cfg["in_channels"], r, r, r, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    ctx = torch.randn(B, 1, cfg["context_dim"], generator=g)
    return x, t, ctx


def vq_input(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    r = cfg["resolution"]
    return (torch.randn(B, 1, r, r, r, generator=g) * 0.1).clamp(-0.2, 0.2)   # dataset clamp, threedfront_dataset.py:391


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}

    # ---- UNet: eps = DiffusionUNet(x, t, c_crossattn=[ctx]) ------------------------------------------
    for tag, cfg, B in (("tiny", D.UNET_TINY, 2), ("full", D.UNET_FULL, 2)):
        m = R.ref_unet(cfg, SEED_UNET)
        x, t, ctx = unet_inputs(cfg, B, 100)
        eps = m(x, t, c_crossattn=[ctx])
        np.savez_compressed(os.path.join(HERE, f"unet_{tag}.npz"), x=x.numpy(), t=t.numpy(), ctx=ctx.numpy(),
                            eps=eps.numpy(), weight_seed=SEED_UNET)
        print(f"unet_{tag}: eps absmax {eps.abs().max():.4f} mean|eps| {eps.abs().mean():.4f}")

        if tag == "tiny":
            # ---- DDIM with classifier-free guidance through the reference's own DDIMSampler ----------
            from model.networks.diffusion_networks.samplers.ddim import DDIMSampler
            sched = D.register_schedule(**D.DIFFUSION)

            class Host:   # the attributes DDIMSampler reads from SDFusionText2ShapeModel
                num_timesteps = 1000
                device = "cpu"
                betas, alphas_cumprod, alphas_cumprod_prev = sched["betas"], sched["alphas_cumprod"], sched["alphas_cumprod_prev"]

                def apply_model(self, x_noisy, tt, cond):   # sdfusion_txt2shape_model.py:275-291
                    return m(x_noisy, tt, c_crossattn=[cond])
            DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)   # keep tensors on CPU
            sampler = DDIMSampler(Host())
            g = torch.Generator().manual_seed(101)
            r = cfg["image_size"]
            xT = torch.randn(1, 3, r, r, r, generator=g).repeat(3, 1, 1, 1, 1)   # shared noise, rel2shape :487-491
            c = torch.randn(3, 1, cfg["context_dim"], generator=g)
            uc = torch.randn(3, 1, cfg["context_dim"], generator=g)
            sampler.make_schedule(ddim_num_steps=100, ddim_eta=0.0, verbose=False)
            steps = np.flip(sampler.ddim_timesteps)
            x_cur, xs, p0s = xT, [], []
            for i in range(4):
                index = len(steps) - i - 1
                ts = torch.full((3,), int(steps[i]), dtype=torch.long)
                x_cur, p0 = sampler.p_sample_ddim(x_cur, c, ts, index=index, unconditional_guidance_scale=3.0,
                                                  unconditional_conditioning=uc)
                xs.append(x_cur.numpy()); p0s.append(p0.numpy())
            np.savez_compressed(os.path.join(HERE, "ddim_tiny.npz"), x_T=xT.numpy(), c=c.numpy(), uc=uc.numpy(),
                                x_steps=np.stack(xs), pred_x0_steps=np.stack(p0s), timesteps=np.ascontiguousarray(steps[:4]),
                                ddim_timesteps=sampler.ddim_timesteps, ddim_alphas=np.asarray(sampler.ddim_alphas),
                                ddim_alphas_prev=np.asarray(sampler.ddim_alphas_prev), weight_seed=SEED_UNET)
            print("ddim_tiny: 4 guided steps, |x| ", float(np.abs(xs[-1]).max()))
        del m

    # ---- VQ-VAE ---------------------------------------------------------------------------------------
    for tag, cfg in (("tiny", V.VQ_TINY), ("full", V.VQ_FULL)):
        m = R.ref_vqvae(cfg, SEED_VQ)
        x = vq_input(cfg, 1, 200)
        z = m(x, forward_no_quant=True, encode_only=True)
        zq, _, (_, _, idx) = m.quantize(z, is_voxel=True)
        dec = m.decode_no_quant(z)
        sub = dec[:, :, ::4, ::4, ::4] if tag == "full" else dec
        np.savez_compressed(os.path.join(HERE, f"vqvae_{tag}.npz"), z=z.numpy(), idx=idx.numpy().astype(np.int32),
                            dec_sub=sub.numpy(), dec_sum=float(dec.double().sum()), dec_abs_sum=float(dec.double().abs().sum()),
                            weight_seed=SEED_VQ, input_seed=200)
        print(f"vqvae_{tag}: z absmax {z.abs().max():.3f} dec absmax {dec.abs().max():.3f} codes used {idx.unique().numel()}")
        del m

    # ---- scene-graph conditioning: encoder_2 + rel_mlp -------------------------------------------------
    for tag, cfg in (("tiny", G.GCN_TINY), ("full", G.GCN_FULL)):
        m = R.RefE2(cfg)
        Wt.fill_module_(m, SEED_GCN)
        inp = R.synth_graph(cfg, 11, 30, seed=300)
        rec = {"z": inp[0].numpy(), "objs": inp[1].numpy(), "triples": inp[2].numpy(), "text": inp[3].numpy(), "rel": inp[4].numpy()}
        for training in (False, True):
            m.train(training)
            uc, c = m(*inp)
            rec[f"uc_{'train' if training else 'eval'}"] = uc.numpy()
            rec[f"c_{'train' if training else 'eval'}"] = c.numpy()
        np.savez_compressed(os.path.join(HERE, f"gcn_{tag}.npz"), weight_seed=SEED_GCN, **rec)
        print(f"gcn_{tag}: c absmax {np.abs(rec['c_eval']).max():.3f}")


if __name__ == "__main__":
    main()
