"""Mirror of pcdet/models/detectors/gd_mae.py:4-37 (GDMAE) and of the part of
Detector3DTemplate that the MAE path needs (pcdet/models/detectors/detector3d_template.py:15-100):
the ``global_step`` buffer, NAME-keyed construction of ``vfe`` and ``backbone_3d`` and the
``forward(batch_dict) -> (ret_dict, tb_dict, disp_dict)`` training contract of
model_func (pcdet/models/__init__.py:29-39)."""
import torch
import torch.nn as nn

from ..backbones_3d import vfe as _vfe_registry
from .. import backbones_3d as _backbone_registry
from .. import backbones_2d as _backbone2d_registry
from .. import dense_heads as _dense_head_registry


class Detector3DTemplate(nn.Module):
    def __init__(self, model_cfg, num_class, dataset, logger=None):
        super().__init__()
        self.model_cfg, self.num_class, self.dataset, self.logger = model_cfg, num_class, dataset, logger
        self.class_names = getattr(dataset, 'class_names', None)
        self.register_buffer('global_step', torch.LongTensor(1).zero_())
        self.module_topology = ['vfe', 'backbone_3d']

    @property
    def mode(self):
        return 'TRAIN' if self.training else 'TEST'

    def update_global_step(self):
        self.global_step += 1

    def build_networks(self):
        info = {'module_list': [],
                'num_rawpoint_features': self.dataset.point_feature_encoder.num_point_features,
                'num_point_features': self.dataset.point_feature_encoder.num_point_features,
                'grid_size': self.dataset.grid_size, 'point_cloud_range': self.dataset.point_cloud_range,
                'voxel_size': self.dataset.voxel_size}
        for name in self.module_topology:
            module, info = getattr(self, 'build_%s' % name)(model_info_dict=info)
            self.add_module(name, module)
        return info['module_list']

    def build_vfe(self, model_info_dict):
        if self.model_cfg.get('VFE', None) is None:
            return None, model_info_dict
        m = _vfe_registry.__all__[self.model_cfg.VFE.NAME](
            model_cfg=self.model_cfg.VFE, num_point_features=model_info_dict['num_rawpoint_features'],
            point_cloud_range=model_info_dict['point_cloud_range'], voxel_size=model_info_dict['voxel_size'],
            grid_size=model_info_dict['grid_size'])
        model_info_dict['num_point_features'] = m.get_output_feature_dim()
        model_info_dict['module_list'].append(m)
        return m, model_info_dict

    def build_backbone_3d(self, model_info_dict):
        if self.model_cfg.get('BACKBONE_3D', None) is None:
            return None, model_info_dict
        m = _backbone_registry.__all__[self.model_cfg.BACKBONE_3D.NAME](
            model_cfg=self.model_cfg.BACKBONE_3D, input_channels=model_info_dict['num_point_features'],
            grid_size=model_info_dict['grid_size'], voxel_size=model_info_dict['voxel_size'],
            point_cloud_range=model_info_dict['point_cloud_range'])
        model_info_dict['module_list'].append(m)
        model_info_dict['num_point_features'] = m.num_point_features
        return m, model_info_dict


    def build_backbone_2d(self, model_info_dict):
        """detector3d_template.py:114-124"""
        if self.model_cfg.get('BACKBONE_2D', None) is None:
            return None, model_info_dict
        m = _backbone2d_registry.__all__[self.model_cfg.BACKBONE_2D.NAME](
            model_cfg=self.model_cfg.BACKBONE_2D, input_channels=model_info_dict.get('num_bev_features', None))
        model_info_dict['module_list'].append(m)
        model_info_dict['num_bev_features'] = m.num_bev_features
        return m, model_info_dict

    def build_dense_head(self, model_info_dict):
        """detector3d_template.py:141-157"""
        if self.model_cfg.get('DENSE_HEAD', None) is None:
            return None, model_info_dict
        m = _dense_head_registry.__all__[self.model_cfg.DENSE_HEAD.NAME](
            model_cfg=self.model_cfg.DENSE_HEAD, input_channels=model_info_dict['num_bev_features'],
            num_class=self.num_class if not self.model_cfg.DENSE_HEAD.CLASS_AGNOSTIC else 1, class_names=self.class_names,
            grid_size=model_info_dict['grid_size'], point_cloud_range=model_info_dict['point_cloud_range'],
            predict_boxes_when_training=self.model_cfg.get('ROI_HEAD', False), voxel_size=model_info_dict.get('voxel_size', False),
            backbone_channels=model_info_dict.get('backbone_channels', None))
        model_info_dict['module_list'].append(m)
        return m, model_info_dict

    # ------------------------------------------------------------------ checkpoint interop (detector3d_template.py:361-442)
    def _load_state_dict(self, model_state_disk, *, strict=True):
        """Copies every tensor of a reference-format ``model_state`` whose key and shape match into this model.  Sparse-conv
        weights stored in another spconv layout are adapted first (detector3d_template.py:361-380): spconv 1.x keeps
        (k.., C_in, C_out); this model keeps spconv 2.x's (C_out, k.., C_in).  The reference handles the
        transposed (k.., C_out, C_in) form and the 5-D (3-D conv) implicit form; the 4-D case of this model's 2-D convs
        (kH, kW, C_in, C_out) -> (C_out, kH, kW, C_in) is the same permutation one dimension lower and is accepted too."""
        from ...utils.spconv_utils import find_all_spconv_keys
        state_dict = self.state_dict()
        spconv_keys = find_all_spconv_keys(self)
        update_model_state = {}
        for key, val in model_state_disk.items():
            if key in spconv_keys and key in state_dict and state_dict[key].shape != val.shape:
                val_native = val.transpose(-1, -2)                 # (k.., c_in, c_out) -> (k.., c_out, c_in)
                if val_native.shape == state_dict[key].shape:
                    val = val_native.contiguous()
                else:
                    assert val.dim() in (4, 5), 'sparse-conv weights are 4-D (2-D conv) or 5-D (3-D conv)'
                    val_implicit = val.permute(val.dim() - 1, *range(val.dim() - 1))   # -> (c_out, k.., c_in)
                    if val_implicit.shape == state_dict[key].shape:
                        val = val_implicit.contiguous()
            if key in state_dict and state_dict[key].shape == val.shape:
                update_model_state[key] = val
        if strict:
            self.load_state_dict(update_model_state)
        else:
            # in-place copies: parameters that live in a trainer's flat bucket stay views of it
            with torch.no_grad():
                for key, val in update_model_state.items():
                    state_dict[key].copy_(val)
        return state_dict, update_model_state

    def load_params_from_file(self, filename, logger=None, to_cpu=False):
        """pre-trained weights only (SSL -> finetune transfer: whatever matches by key and shape), :393-412"""
        import os
        if not os.path.isfile(filename):
            raise FileNotFoundError
        loc_type = torch.device('cpu') if to_cpu else None
        checkpoint = torch.load(filename, map_location=loc_type, weights_only=False)
        state_dict, update_model_state = self._load_state_dict(checkpoint['model_state'], strict=False)
        if logger is not None:
            for key in state_dict:
                if key not in update_model_state:
                    logger.info('Not updated weight %s: %s' % (key, str(state_dict[key].shape)))
            logger.info('==> Done (loaded %d/%d)' % (len(update_model_state), len(state_dict)))
        return len(update_model_state), len(state_dict)

    def load_params_with_optimizer(self, filename, to_cpu=False, optimizer=None, logger=None):
        """resume: strict weights + optimizer state (+ the `_optim` side file), -> (it, epoch), :414-442.  ``optimizer`` is a
        MAETrainer (``load_state_dict`` takes the reference's torch-Adam ``optimizer_state`` as well as its own format)."""
        import os
        if not os.path.isfile(filename):
            raise FileNotFoundError
        loc_type = torch.device('cpu') if to_cpu else None
        checkpoint = torch.load(filename, map_location=loc_type, weights_only=False)
        epoch, it = checkpoint.get('epoch', -1), checkpoint.get('it', 0.0)
        self._load_state_dict(checkpoint['model_state'], strict=True)
        if optimizer is not None:
            if checkpoint.get('optimizer_state', None) is not None:
                optimizer.load_state_dict(checkpoint['optimizer_state'])
            else:
                assert filename[-4] == '.', filename
                optimizer_filename = '%s_optim.%s' % (filename[:-4], filename[-3:])
                if os.path.exists(optimizer_filename):
                    optimizer.load_state_dict(torch.load(optimizer_filename, map_location=loc_type, weights_only=False)['optimizer_state'])
            if hasattr(optimizer, 'it'):
                optimizer.it = int(it)
        if logger is not None:
            logger.info('==> Done')
        return it, epoch


class GDMAE(Detector3DTemplate):
    def __init__(self, model_cfg, num_class, dataset, logger=None):
        super().__init__(model_cfg=model_cfg, num_class=num_class, dataset=dataset, logger=logger)
        self.module_list = self.build_networks()

    # ------------------------------------------------------------------ index pipeline one step ahead
    def prefetch_index(self, batch_dict, ready_event=None):
        """Runs the parameter-free half of the forward pass for ``batch_dict`` - dynamic voxelisation, MAE mask, visible
        sites, pyramid site sets, neighbour maps, window tables, SRA work units - on a side stream, typically for the NEXT
        batch while this step's backward is still queued (MAETrainer.step(batch, next_batch=...)).  Its two count reads then
        wait for the side stream only, and the later forward(batch_dict) has no host sync and no table building left: the
        main stream never drains between steps.  ``ready_event``: recorded when batch_dict['points'] is on the device."""
        from ..backbones_3d.spt_backbone_mae import SPTBackboneMAE
        from ..backbones_3d.vfe.dyn_vfe import DynVFE
        mods = self.module_list
        if not (len(mods) >= 2 and isinstance(mods[0], DynVFE) and isinstance(mods[1], SPTBackboneMAE)):
            return batch_dict
        import torch
        dev = batch_dict['points'].device
        if getattr(self, '_side_stream', None) is None:
            self._side_stream = torch.cuda.Stream(device=dev, priority=-1)
        side = self._side_stream
        if ready_event is not None:
            side.wait_event(ready_event)
        batch_dict['points'].record_stream(side)     # allocated on the caller's stream, read (and released) here
        with torch.cuda.stream(side):
            mods[0].index_pass(batch_dict)
            mods[1].index_pass(batch_dict)
            ev = torch.cuda.Event()
            ev.record(side)
        batch_dict['_index_event'] = ev
        return batch_dict

    @staticmethod
    def _record_stream(obj, stream, seen):
        """every tensor reachable from obj was allocated on the side stream and is about to be used on ``stream``"""
        import torch
        if obj is None or isinstance(obj, (int, float, str, bool)) or id(obj) in seen:
            return
        seen.add(id(obj))
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                obj.record_stream(stream)
        elif isinstance(obj, dict):
            for v in obj.values():
                GDMAE._record_stream(v, stream, seen)
        elif isinstance(obj, (list, tuple)):
            for v in obj:
                GDMAE._record_stream(v, stream, seen)
        elif hasattr(obj, '__dict__') or hasattr(obj, '__slots__'):
            names = list(getattr(obj, '__dict__', {}).keys()) + list(getattr(type(obj), '__slots__', ()))
            for n in names:
                try:
                    v = object.__getattribute__(obj, n)
                except AttributeError:
                    continue
                GDMAE._record_stream(v, stream, seen)

    def forward(self, batch_dict):
        ev = batch_dict.pop('_index_event', None)
        if ev is not None:
            import torch
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)                       # the prefetched index structures are complete
            self._record_stream(batch_dict, cur, set())
        # DynVFE followed by SPTBackboneMAE: let the backbone schedule the VFE's feature pass after its own index kernels
        # (see DynVFE.forward); any other module order runs every module to completion like the reference
        from ..backbones_3d.spt_backbone_mae import SPTBackboneMAE
        from ..backbones_3d.vfe.dyn_vfe import DynVFE
        mods = self.module_list
        if len(mods) >= 2 and isinstance(mods[0], DynVFE) and isinstance(mods[1], SPTBackboneMAE):
            batch_dict['defer_vfe_features'] = True
        for cur_module in mods:
            batch_dict = cur_module(batch_dict)
        if self.training:
            loss, tb_dict, disp_dict = self.get_training_loss()
            return {'loss': loss}, tb_dict, disp_dict
        return self.post_processing(batch_dict)

    def post_processing(self, batch_dict):
        return {}, {}

    def get_training_loss(self, sync_for_logging=False):
        """gd_mae.py:27-37.  The reference puts ``loss_rpn.item()`` (a host sync) into tb_dict every
        step; here the tensor itself is stored unless ``sync_for_logging`` is set."""
        loss_rpn, tb_dict = self.backbone_3d.get_loss()
        tb_dict = {'loss_rpn': loss_rpn.item() if sync_for_logging else loss_rpn.detach(), **tb_dict}
        return loss_rpn, tb_dict, {}
