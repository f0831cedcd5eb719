import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import torch_sd
from oracle import trpx_oracle as O
from trphysx_b200.embedding.embedding_cylinder import CylinderEmbeddingTrainer
dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
for (B, T, wr, wd, wg) in ((3, 2, 10, 10, 1e-2), (3, 2, 10, 0, 0), (3, 2, 0, 10, 0), (3, 2, 0, 0, 1e-2), (5, 2, 10, 10, 1e-2), (3, 1, 10, 10, 1e-2), (6, 1, 10, 0, 0), (9, 1, 10, 0, 0)):
    head = CylinderEmbeddingTrainer(cfg)
    head.W_REC, head.W_DYN, head.W_REG = float(wr), float(wd), float(wg)
    head.embedding_model.load_state_dict(torch_sd(synth.cylinder_embedding_weights(cfg, 612)))
    head.to(dev)
    sd = O.to_torch(synth.cylinder_embedding_weights(cfg, 612))
    x, visc = synth.cylinder_fields(B * T, seed=77)
    states_c = torch.from_numpy(x).view(B, T, 3, 64, 128)
    visc_c = torch.from_numpy(visc).view(B, T, 1)[:, 0]
    ref_loss, ref_rec, ref_grad = O.cylinder_train_grads(sd, cfg, states_c, visc_c, w_rec=wr, w_dyn=wd, w_reg=wg)
    loss, rec = head(states_c.to(dev), visc_c.to(dev))
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()]).cpu()
    bad = []
    off = 0
    for k, p in head.embedding_model.named_parameters():
        n = p.numel()
        if float(ref_grad[off:off + n].norm()) > 0:
            e = O.rel_l2(flat[off:off + n], ref_grad[off:off + n])
            if e > 1e-5:
                bad.append(f"{k}:{e:.1e}")
        off += n
    print(f"B={B} T={T} w=({wr},{wd},{wg}): loss {float(loss):.6f} ref {float(ref_loss):.6f}; grad rel-L2 {O.rel_l2(flat, ref_grad):.3e}; bad: {bad[:6]}")
