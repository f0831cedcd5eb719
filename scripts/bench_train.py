"""BASELINE config 3: Lorenz Koopman embedding training step (encoder/decoder + banded Koopman loss),
data-parallel with an NCCL all-reduce of the flat gradient, states [512, 16, 3] per GPU (weak scaling).
Run with torchrun for N > 1.  Prints one JSON line on rank 0.  GPU only."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from tests import synth
from oracle import trpx_oracle as O  # noqa: E402
from tests.util import torch_sd  # noqa: E402


def main():
    from trphysx_b200.config import LorenzConfig
    from trphysx_b200.embedding import LorenzEmbeddingTrainer
    from trphysx_b200.optim import FusedClipAdam, FusedKoopmanStep
    from trphysx_b200.parallel import average_flat_grad
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg_ns = synth.make_config("lorenz")
    head = LorenzEmbeddingTrainer(LorenzConfig())
    head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, 77)))
    head.to(dev)
    B, T = 512, 16
    states = torch.from_numpy(synth.lorenz_trajectories(B, T - 1, seed=100 + rank)).to(dev)
    opt = FusedClipAdam(head.parameters(), lr=1e-3, weight_decay=1e-8, max_norm=0.1)

    def step():
        loss, loss_rec = head(states)
        loss.backward()
        average_flat_grad(opt.flat_grad(), world)
        opt.step()
        opt.zero_grad()
        return loss

    fused = FusedKoopmanStep(head, opt, world)
    if os.environ.get("TRPX_TRAIN_FUSED", "1") == "1":
        def step():                      # noqa: F811  (no autograd graph: 3 native calls per step)
            return fused(states)[0]

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(K):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K * 1e3
    ms = a.elapsed_time(b) / K
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"workload": "lorenz_koopman_train_step states=[512,16,3]/GPU fwd+bwd+allreduce+clip+adam",
                "n_gpus": world, "ms_per_step": round(float(t), 4), "wall_ms_per_step": round(wall, 4),
                "samples_per_s": round(world * B / float(t) * 1e3, 1), "loss": float(loss)}
        if world == 1:
            sd = O.to_torch(synth.lorenz_embedding_weights(cfg_ns, 77))
            torch.set_num_threads(os.cpu_count())
            cs = states.cpu()
            O.lorenz_train_grads(sd, cfg_ns, cs)
            t1 = time.perf_counter()
            for _ in range(3):
                O.lorenz_train_grads(sd, cfg_ns, cs)
            line["cpu_baseline_ms_per_step_fwd_bwd"] = round((time.perf_counter() - t1) / 3 * 1e3, 2)
            line["cpu_cores"] = os.cpu_count()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
