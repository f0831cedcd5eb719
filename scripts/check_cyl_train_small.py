import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import torch_sd
from trphysx_b200.embedding.embedding_cylinder import CylinderEmbeddingTrainer
dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
B, T = 3, 2
head = CylinderEmbeddingTrainer(cfg)
head.embedding_model.load_state_dict(torch_sd(synth.cylinder_embedding_weights(cfg, 612)))
head.to(dev)
x, visc = synth.cylinder_fields(B * T, seed=77)
loss, rec = head(torch.from_numpy(x).view(B, T, 3, 64, 128).to(dev), torch.from_numpy(visc).view(B, T, 1)[:, 0].to(dev))
loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
