"""Per-parameter-group error of the Koopman training step gradient (config 3 shape) for the tensor-core and the FP32
decoder kernels against the oracle in float32 and float64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import torch_sd
from oracle import trpx_oracle as O
from trphysx_b200.config import PhysConfig
from trphysx_b200.embedding import LorenzEmbeddingTrainer

dev = torch.device("cuda:0")
cfg_ns = synth.make_config("lorenz")
states = torch.from_numpy(synth.lorenz_trajectories(512, 15, seed=77))
sd = O.to_torch(synth.lorenz_embedding_weights(cfg_ns, 77))
l32, _, g32 = O.lorenz_train_grads(sd, cfg_ns, states)
sd64 = O.to_torch(synth.lorenz_embedding_weights(cfg_ns, 77), torch.float64)
l64, _, g64 = O.lorenz_train_grads(sd64, cfg_ns, states.double())
res = {}
for env in ("1", "0"):
    os.environ["TRPX_KOOPMAN_TC"] = "0" if env == "1" else "1"
    head = LorenzEmbeddingTrainer(PhysConfig(**vars(cfg_ns)))
    head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, 77)))
    head.to(dev)
    loss, _ = head(states.to(dev))
    loss.backward()
    names = [n for n, _ in head.named_parameters()]
    grads = [p.grad.reshape(-1).cpu().double() for p in head.parameters()]
    res[env] = (float(loss), names, grads)
rel = lambda a, b: float((a - b).norm() / b.norm())
print("loss: oracle32 %.6f oracle64 %.6f simt %.6f tc %.6f" % (float(l32), float(l64), res["1"][0], res["0"][0]))
off = 0
flat = {k: torch.cat(v[2]) for k, v in res.items()}
print("whole gradient vs f64: oracle32 %.2e  simt %.2e  tc %.2e" % (rel(g32.double(), g64), rel(flat["1"], g64), rel(flat["0"], g64)))
for i, n in enumerate(res["0"][1]):
    k = res["0"][2][i].numel()
    ref = g64[off:off + k]
    print(f"{n:42s} norm {float(ref.norm()):10.3e}  oracle32 {rel(g32[off:off+k].double(), ref):.2e}  simt {rel(res['1'][2][i], ref):.2e}  tc {rel(res['0'][2][i], ref):.2e}")
    off += k
