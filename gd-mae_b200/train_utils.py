"""Checkpoint I/O of the training loop in the reference's file format (tools/train_utils/train_utils.py:140-174):
``{'epoch', 'it', 'model_state', 'optimizer_state', 'version'}`` written with torch.save as ``<name>.pth``.  A file written
here loads in the reference (``load_params_from_file`` / ``load_params_with_optimizer``) and the other way round;
``model_state`` uses the reference's state_dict keys and spconv-2.x weight layout."""
import torch

VERSION = 'gd-mae_b200+r2'


def model_state_to_cpu(model_state):
    return type(model_state)((k, v.cpu()) for k, v in model_state.items())


def checkpoint_state(model=None, optimizer=None, epoch=None, it=None):
    """``optimizer``: a MAETrainer; its Adam moments are stored in the reference's ``optimizer_state`` layout"""
    if optimizer is not None and hasattr(optimizer, 'sync_buffers'):
        optimizer.sync_buffers()            # every rank saves rank 0's BatchNorm statistics (what DDP's broadcast_buffers leaves)
    optim_state = optimizer.state_dict(reference_format=True) if optimizer is not None else None
    model_state = model_state_to_cpu(model.state_dict()) if model is not None else None
    return {'epoch': epoch, 'it': it, 'model_state': model_state, 'optimizer_state': optim_state, 'version': VERSION}


def save_checkpoint(state, filename='checkpoint'):
    filename = '{}.pth'.format(filename)
    torch.save(state, filename)
    return filename
