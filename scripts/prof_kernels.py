"""Driver for ncu: launches every non-rollout kernel of the path once at a representative size.  GPU only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch
from tests import synth
from tests.util import make_cylinder, make_gpt2, make_lorenz, torch_sd
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
with torch.no_grad():
    cfg = synth.make_config("lorenz")
    emb = make_lorenz(cfg, 1, "lorenz", dev)
    x = torch.randn(2_000_000, 3, device=dev)
    tr = make_gpt2(cfg, 2, False, dev)
    xs = torch.randn(4096, 64, 32, device=dev)
    ccfg = synth.make_config("cylinder")
    cyl = make_cylinder(ccfg, 3, dev)
    cx, cv = synth.cylinder_fields(64, seed=0)
    cx, cv = torch.from_numpy(cx).to(dev), torch.from_numpy(cv).to(dev)
    for _ in range(reps):
        g = emb.embed(x)
        emb.recover(g)
        emb.koopmanOperation(g)
        tr(xs, use_cache=True)
        cg = cyl.embed(cx, cv)
        cyl.recover(cg)
        cyl.koopmanOperation(cg, cv)
        torch.cuda.synchronize()
from trphysx_b200.config import LorenzConfig
from trphysx_b200.embedding import LorenzEmbeddingTrainer
from trphysx_b200.optim import FusedClipAdam
head = LorenzEmbeddingTrainer(LorenzConfig())
head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg, 77)))
head.to(dev)
states = torch.from_numpy(synth.lorenz_trajectories(512, 15, seed=1)).to(dev)
opt = FusedClipAdam(head.parameters(), lr=1e-3, weight_decay=1e-8, max_norm=0.1)
for _ in range(reps):
    loss, _ = head(states)
    loss.backward()
    opt.step()
    opt.zero_grad()
torch.cuda.synchronize()
print("ok")
