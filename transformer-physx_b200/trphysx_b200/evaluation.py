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


@torch.no_grad()
def rollout_to_host(embedding_model, transformer, x0_host: Tensor, max_length: int, states_host: Tensor,
                    device=None, use_cache: bool = True, chunks: int = 4, copy_stream=None,
                    visc_host: Tensor = None, pipeline: bool = True) -> Tensor:
    """The evaluation pipeline of `Trainer.eval_states` around `generate()` with HOST buffers:
    x0_host [B, *state_dims] (pinned; plus visc_host [B, 1] for the Cylinder embedding) -> device -> embed ->
    generate(max_length) -> recover -> states_host
    [B, max_length, *state_dims] (pinned).  `recover` runs in `chunks` slices of trajectories and each slice's states
    are copied to the host on `copy_stream` while the next slice is recovered; with `pipeline` (default) large batches
    of the E = 32 models are also DECODED slice by slice (one round of resident trajectories each), so that the copies
    overlap the decoding and only the last slice's copy is exposed.  Returns
    `states_host` once the current stream has been made to wait for the copies (call torch.cuda.synchronize() or
    record an event before reading it on the host)."""
    dev = torch.device(device) if device is not None else embedding_model.devices[0]
    B = x0_host.size(0)
    E = transformer.config.n_embd
    cur = torch.cuda.current_stream(dev)
    if copy_stream is None:
        copy_stream = torch.cuda.Stream(device=dev)
    x = x0_host.to(dev, non_blocking=True)
    if visc_host is not None:           # Cylinder: embed(x, visc) (embedding_cylinder.py:138)
        g0 = embedding_model.embed(x, visc_host.to(dev, non_blocking=True))
    else:
        g0 = embedding_model.embed(x)
    # Large batches of the E = 32 models: the rollout kernel works through the batch in rounds of R resident trajectories
    # anyway, so generate() is called per slice of at most R trajectories (same number of rounds, measured 98.7 against
    # 98.6 ms for 65,536 x 256) and a slice's recover + device->host copy overlap the decoding of the next slice: only the
    # last slice's copy is exposed instead of the whole batch's.
    R = transformer.rollout_round_trajectories(dev, use_cache) if hasattr(transformer, "rollout_round_trajectories") else None
    pipelined = pipeline and R is not None and B >= 2 * R
    if pipelined:
        chunks = (B + R - 1) // R
        out = None
    else:
        out = transformer.generate(g0.unsqueeze(1), max_length=max_length, use_cache=use_cache, return_presents=False)[0]
        chunks = max(1, min(chunks, B))
    for i in range(chunks):
        lo, hi = B * i // chunks, B * (i + 1) // chunks
        if pipelined:
            out_i = transformer.generate(g0[lo:hi].unsqueeze(1), max_length=max_length, use_cache=use_cache,
                                         return_presents=False)[0]
        else:
            out_i = out[lo:hi]
        st = embedding_model.recover(out_i.reshape((hi - lo) * max_length, E))
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            states_host[lo:hi].copy_(st.view([hi - lo, max_length] + list(embedding_model.input_dims)), non_blocking=True)
            st.record_stream(copy_stream)
    cur.wait_stream(copy_stream)
    return states_host
