import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch, time
from tests import synth
from tests.util import torch_sd
from oracle import trpx_oracle as O
from trphysx_b200.embedding.embedding_cylinder import CylinderEmbeddingTrainer
dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
for (B, T) in ((3, 2), (2, 2), (4, 2), (3, 3), (1, 4)):
    head = CylinderEmbeddingTrainer(cfg)
    head.embedding_model.load_state_dict(torch_sd(synth.cylinder_embedding_weights(cfg, 612)))
    head.to(dev)
    sd = O.to_torch(synth.cylinder_embedding_weights(cfg, 612))
    x, visc = synth.cylinder_fields(B * T, seed=77)
    states_c = torch.from_numpy(x).view(B, T, 3, 64, 128)
    visc_c = torch.from_numpy(visc).view(B, T, 1)[:, 0]
    ref_loss, ref_rec, ref_grad = O.cylinder_train_grads(sd, cfg, states_c, visc_c)
    loss, rec = head(states_c.to(dev), visc_c.to(dev))
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()]).cpu()
    print(f"B={B} T={T}: loss {float(loss):.6f} ref {float(ref_loss):.6f}; rec {float(rec):.6f} ref {float(ref_rec):.6f}; grad rel-L2 {O.rel_l2(flat, ref_grad):.3e}")
    off = 0
    for k, p in head.embedding_model.named_parameters():
        n = p.numel()
        e = O.rel_l2(flat[off:off + n], ref_grad[off:off + n])
        if e > 1e-5:
            print(f"    {k}: {e:.3e}")
        off += n
if len(sys.argv) > 1:
    B, T = 16, 4
    head = CylinderEmbeddingTrainer(cfg); head.to(dev)
    x, visc = synth.cylinder_fields(B * T, seed=5)
    st = torch.from_numpy(x).view(B, T, 3, 64, 128).to(dev); v = torch.from_numpy(visc).view(B, T, 1)[:, 0].to(dev)
    for _ in range(2):
        head.zero_grad(); l, _ = head(st, v); l.backward()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        head.zero_grad(); l, _ = head(st, v); l.backward()
    torch.cuda.synchronize(); print(f"B={B} T={T}: {(time.time()-t0)/3*1e3:.2f} ms per fwd+bwd")
