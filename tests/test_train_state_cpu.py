"""CPU suite: checkpoint / resume of the native training step (the optimizer half): the flat AdamW
moments export to and import from torch.optim.AdamW's own state-dict layout — the format of the 'opt' entry the reference
writes (sdfusion_txt2shape_model.py:636-650, VAE.py:334-340) — and flat re-homing keeps the reference's state-dict keys.
No kernel runs here (the step itself is covered on the GPU)."""
import torch

from oracle import denoiser as D


class _Stub:
    num_timesteps = 1000

    def __init__(self, df):
        self.df = df


def _tiny():
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    torch.manual_seed(0)
    return DiffusionUNet(dict(D.UNET_TINY, use_spatial_transformer=True, use_checkpoint=True, legacy=False), conditioning_key="crossattn")


def test_flat_state_round_trips_through_torch_adamw():
    from commonscenes_b200.train import DenoiserTrainStep
    df = _tiny()
    keys_before = {k: tuple(v.shape) for k, v in df.state_dict().items()}
    step = DenoiserTrainStep(_Stub(df), lr=2e-4)
    assert {k: tuple(v.shape) for k, v in df.state_dict().items()} == keys_before        # re-homing keeps keys and shapes
    assert all(p.data_ptr() >= step.flat_p.data_ptr() for p in df.parameters())         # parameters are views of the flat buffer
    g = torch.Generator().manual_seed(1)
    step.flat_m.copy_(torch.randn(step.flat_m.shape, generator=g))
    step.flat_v.copy_(torch.rand(step.flat_v.shape, generator=g))
    step.step_count = 7
    sd = step.optimizer_state_dict()
    # torch's own optimizer accepts it ...
    opt = torch.optim.AdamW(df.parameters(), lr=1e-4)
    opt.load_state_dict(sd)
    params = [p for p in df.parameters() if p.requires_grad]
    assert opt.param_groups[0]["lr"] == 2e-4 and len(opt.state) == len(params)
    for p in params:
        off, n = step.offsets[p], p.numel()
        assert torch.equal(opt.state[p]["exp_avg"].flatten(), step.flat_m[off:off + n])
        assert torch.equal(opt.state[p]["exp_avg_sq"].flatten(), step.flat_v[off:off + n])
        assert float(opt.state[p]["step"]) == 7.0
    # ... and torch's state dict loads back into a fresh native step
    df2 = _tiny()
    step2 = DenoiserTrainStep(_Stub(df2))
    step2.load_optimizer_state_dict(opt.state_dict())
    assert step2.step_count == 7 and int(step2.step_dev) == 7 and step2.lr == 2e-4
    for p, p2 in zip(step.params, step2.params):          # (alignment gaps between parameters carry no state)
        a, b, n = step.offsets[p], step2.offsets[p2], p.numel()
        assert torch.equal(step2.flat_m[b:b + n], step.flat_m[a:a + n]) and torch.equal(step2.flat_v[b:b + n], step.flat_v[a:a + n])


def test_torch_adamw_state_after_real_steps_imports():
    """A state dict produced by torch.optim.AdamW.step() itself (what a reference checkpoint holds)."""
    from commonscenes_b200.train import DenoiserTrainStep
    df = _tiny()
    opt = torch.optim.AdamW(df.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(2)
    for _ in range(3):
        for p in df.parameters():
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
        opt.step()
    sd = opt.state_dict()
    df2 = _tiny()
    df2.load_state_dict(df.state_dict())
    step = DenoiserTrainStep(_Stub(df2))
    step.load_optimizer_state_dict(sd)
    assert step.step_count == 3
    for i, p in enumerate(step.params):
        off, n = step.offsets[p], p.numel()
        assert torch.equal(step.flat_m[off:off + n].view(p.shape), sd["state"][i]["exp_avg"])
    for (k, a), (_, b) in zip(df.state_dict().items(), df2.state_dict().items()):
        assert torch.equal(a, b), k
