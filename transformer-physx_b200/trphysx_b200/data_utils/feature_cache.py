"""The reference's pickle feature cache, read and written as is.

`PhysicalDataset` (trphysx/data_utils/dataset_phys.py:36-111) embeds every time series of an HDF5 file once, cuts the
embedded series into blocks and pickles `(examples, states)` — two lists of CPU float32 tensors, `examples[i]`
`[block_size, n_embd]`, `states[i]` `[block_size, *state_dims]` (filled only for `eval=True`) — to
`cached{ndata}_{EmbedderClass}_{block_size}_{data file name}` next to the data (or in `cache_path`), under a
`FileLock` so that only the first process of a distributed job embeds.  This module keeps that file format, name and
locking, so caches written by the reference load here and vice versa; the embedding itself is ONE bulk call per series
on the device (`evaluation.embed_series`) instead of the reference's per-series `.cpu()` loop.

Reading the HDF5 container is out of scope (h5py is not part of this path): the series are passed in as arrays /
tensors, e.g. `{key: h5_file[key][...]}`.
"""
import logging
import os
import pickle
import time
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch
from torch import Tensor

logger = logging.getLogger(__name__)
SeriesMap = Union[Dict[str, object], Iterable[Tuple[str, object]]]


def cached_features_file(file_path: str, embedder, block_size: int, ndata: int = -1, cache_path: Optional[str] = None) -> str:
    """dataset_phys.py:60-65: `<cache_path or data dir>/cached{ndata}_{class name}_{block_size}_{file name}`."""
    directory, filename = os.path.split(file_path)
    if cache_path is None or not os.path.isdir(cache_path):
        cache_path = directory
    return os.path.join(cache_path, "cached{}_{}_{}_{}".format(ndata, embedder.__class__.__name__, str(block_size), filename))


def read_cache(cached_file: str) -> Tuple[List[Tensor], List[Tensor]]:
    """dataset_phys.py:88-99."""
    assert os.path.isfile(cached_file), 'Provided cache file path does not exist!'
    start = time.time()
    with open(cached_file, "rb") as handle:
        examples, states = pickle.load(handle)
    logger.info("Loading features from cached file %s [took %.3f s]", cached_file, time.time() - start)
    return examples, states


def write_cache(cached_file: str, examples: List[Tensor], states: List[Tensor]) -> None:
    """dataset_phys.py:101-111: pickle.HIGHEST_PROTOCOL of the tuple (examples, states), CPU tensors."""
    start = time.time()
    os.makedirs(os.path.dirname(os.path.abspath(cached_file)), exist_ok=True)
    with open(cached_file, "wb") as handle:
        pickle.dump(([e.cpu() for e in examples], [s.cpu() for s in states]), handle, protocol=pickle.HIGHEST_PROTOCOL)
    logger.info("Saving features into cached file %s [took %.3f s]", cached_file, time.time() - start)


@torch.no_grad()
def embed_series_blocks(embedder, series: SeriesMap, block_size: int, stride: int = 1, ndata: int = -1, eval: bool = False,
                        visc_of_key=None) -> Tuple[List[Tensor], List[Tensor]]:
    """`LorenzDataset.embed_data` / `CylinderDataset.embed_data` (dataset_lorenz.py:22-47, dataset_cylinder.py:21-52):
    every series `[T, *state_dims]` is embedded in one call on the embedder's device and cut into blocks of
    `block_size` every `stride` steps; `ndata > 0` stops after that many series.  `visc_of_key(key) -> float` supplies
    the viscosity of a Cylinder series (the reference uses `2.0 / float(key)`, dataset_cylinder.py:36)."""
    from ..evaluation import embed_series
    dev = embedder.devices[0]
    items = series.items() if hasattr(series, "items") else series
    examples, states = [], []
    samples = 0
    embedder.eval()
    for key, data in items:
        data_series = torch.as_tensor(data, dtype=torch.float32).to(dev).reshape([-1] + list(embedder.input_dims))
        extra = []
        if visc_of_key is not None:
            extra = [float(visc_of_key(key)) * torch.ones(1, data_series.size(0), 1, device=dev)]
        embedded = embed_series(embedder, data_series.unsqueeze(0), *extra)[0].cpu()
        cpu_series = data_series.cpu() if eval else None
        for i in range(0, data_series.size(0) - block_size + 1, stride):
            examples.append(embedded[i: i + block_size])
            if eval:
                states.append(cpu_series[i: i + block_size])
        samples += 1
        if 0 < ndata <= samples:
            break
    logger.info('Collected %d time-series. Total of %d blocks.', samples, len(examples))
    return examples, states


def build_feature_cache(embedder, file_path: str, series: Optional[SeriesMap], block_size: int, stride: int = 1, ndata: int = -1,
                        eval: bool = False, overwrite_cache: bool = False, cache_path: Optional[str] = None,
                        visc_of_key=None) -> Tuple[List[Tensor], List[Tensor]]:
    """The constructor logic of `PhysicalDataset` (dataset_phys.py:53-86): under the cache file's lock, load the cache
    if it exists (and `overwrite_cache` is False), otherwise embed `series` and write it.  `file_path` only names the
    cache (it is the HDF5 path in the reference); `series` may be None when the cache is expected to exist."""
    from filelock import FileLock
    cached = cached_features_file(file_path, embedder, block_size, ndata, cache_path)
    with FileLock(cached + ".lock"):
        if os.path.exists(cached) and not overwrite_cache:
            return read_cache(cached)
        if series is None:
            raise FileNotFoundError(f"no feature cache at {cached} and no series to embed")
        examples, states = embed_series_blocks(embedder, series, block_size, stride, ndata, eval, visc_of_key)
        write_cache(cached, examples, states)
        return examples, states
