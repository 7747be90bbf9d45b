import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from test_shape_branch_train_gpu import _scene_batch, _model
from commonscenes_b200.train import ShapeBranchTrainStep
b = {k: v.cuda() for k, v in _scene_batch(2, 4, 6, 36, 16, seed=11).items()}
O, T = b["objs"].shape[0], b["triples"].shape[0]
g = torch.Generator().manual_seed(9)
t = torch.randint(0, 1000, (O,), generator=g).cuda(); noise = torch.randn(O, 3, 16, 16, 16, generator=g).cuda()
def rel(a, c): return float((a - c).norm() / (a.norm() + 1e-30))
me1, me2, mg = _model(55), _model(55), _model(55)
e1, e2, gr = ShapeBranchTrainStep(me1), ShapeBranchTrainStep(me2), ShapeBranchTrainStep(mg)
gr.capture(O, T)
args = (b["z"], b["objs"], b["triples"], b["text"], b["rel"], b["sdfs"])
l1, dz1 = e1.step(*args, t=t, noise=noise); dz1 = dz1.clone()
l2, dz2 = e2.step(*args, t=t, noise=noise); dz2 = dz2.clone()
lg, dzg = gr.step_graphed(*args, t=t, noise=noise); dzg = dzg.clone()
print("loss", float(l1), float(l2), float(lg))
print("d_z eager1 vs eager2", rel(dz1, dz2), " eager1 vs graphed", rel(dz1, dzg))
print("graph-side flat_g  e1 vs e2", rel(e1.graph_params.flat_g, e2.graph_params.flat_g), " e1 vs graphed", rel(e1.graph_params.flat_g, gr.graph_params.flat_g))
print("denoiser flat_g    e1 vs e2", rel(e1.denoiser.flat_g, e2.denoiser.flat_g), " e1 vs graphed", rel(e1.denoiser.flat_g, gr.denoiser.flat_g))
print("graph-side flat_p  e1 vs e2", rel(e1.graph_params.flat_p, e2.graph_params.flat_p), " e1 vs graphed", rel(e1.graph_params.flat_p, gr.graph_params.flat_p))
# second replay with the same inputs on a fresh graphed model? replay again: params changed, so just check finiteness
lg2, dzg2 = gr.step_graphed(*args, t=t, noise=noise)
l1b, dz1b = e1.step(*args, t=t, noise=noise)
print("step 2 loss eager", float(l1b), "graphed", float(lg2), " d_z rel", rel(dz1b, dzg2))
