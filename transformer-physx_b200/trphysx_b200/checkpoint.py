"""Checkpoint files shared by the embedding and transformer modules.

The on-disk convention is the reference's (`{model_name}{epoch}.pth` holding a plain state_dict, written into a
directory; a file path or a directory is accepted when loading — trphysx/embedding/embedding_model.py:81-119,
trphysx/transformer/phys_transformer_base.py:99-137), so checkpoints travel in both directions."""
import os

import torch


def checkpoint_path(directory: str, model_name: str, epoch: int) -> str:
    return os.path.join(directory, "{}{:d}.pth".format(model_name, int(epoch)))


def write_state_dict(module: torch.nn.Module, directory: str, epoch: int) -> str:
    """Save `module.state_dict()` as `<directory>/<model_name><epoch>.pth`; `directory` must not be a file."""
    if os.path.isfile(directory):
        raise FileNotFoundError("expected a directory to save into, got the file {}".format(directory))
    os.makedirs(directory, exist_ok=True)
    target = checkpoint_path(directory, module.model_name, epoch)
    torch.save(module.state_dict(), target)
    return target


def read_state_dict(module: torch.nn.Module, file_or_directory: str, epoch: int) -> str:
    """Load a state_dict from a `.pth` file, or from `<directory>/<model_name><epoch>.pth`, onto the CPU first
    (parameters keep their current device), strictly."""
    if os.path.isdir(file_or_directory):
        source = checkpoint_path(file_or_directory, module.model_name, epoch)
    elif os.path.isfile(file_or_directory):
        source = file_or_directory
    else:
        raise FileNotFoundError("no checkpoint file or directory at {}".format(file_or_directory))
    module.load_state_dict(torch.load(source, map_location="cpu"))
    return source
