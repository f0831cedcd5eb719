"""Steps either side of the rollout.

`eval_states` is the computation of `Trainer.eval_states` (trphysx/utils/trainer.py:353-390) without its plotting:
recover the predicted embeddings to state space and return the MSE against the target states.  For the Lorenz /
Rossler embeddings the recovery and the error are ONE launch and the recovered states are never written
(`trpx_mlpemb_recover_mse`); the Cylinder decoder writes its frames (they are the product there) and the error is
a device-side reduction.  `embed_series` is the bulk embedding of `LorenzDataset.embed_data`
(trphysx/data_utils/dataset_lorenz.py:22-47): a whole series tensor in one call, on the device, no `.cpu()`."""
import torch
from torch import Tensor


@torch.no_grad()
def eval_states(embedding_model, pred_embeds: Tensor, states: Tensor, return_states: bool = False):
    """pred_embeds [B, T, n_embd], states [B, T, *state_dims] -> state MSE (device scalar) [, recovered states]."""
    B, T = pred_embeds.size(0), pred_embeds.size(1)
    g = pred_embeds.contiguous().view(-1, pred_embeds.size(-1))
    states = states.to(g.device)
    if hasattr(embedding_model, "recover_mse"):
        res = embedding_model.recover_mse(g, states.reshape(g.size(0), -1), return_states=return_states)
        if return_states:
            return res[0], res[1].view([B, T] + list(embedding_model.input_dims))
        return res
    out = embedding_model.recover(g).view([B, T] + list(embedding_model.input_dims))
    err = torch.nn.functional.mse_loss(out, states)
    return (err, out) if return_states else err


@torch.no_grad()
def embed_series(embedding_model, series: Tensor, *extra) -> Tensor:
    """series [B, T, *state_dims] (on the model's device) -> embeddings [B, T, n_embd] in one call."""
    B, T = series.size(0), series.size(1)
    flat = series.reshape([B * T] + list(series.shape[2:]))
    args = [e.reshape([B * T] + list(e.shape[2:])) for e in extra]
    return embedding_model.embed(flat, *args).view(B, T, -1)
