"""Mirror of pcdet/models/detectors/centerpoint.py:4-50 (CenterPoint), the detector of the finetune configs
(tools/cfgs/*/gd_mae_iou.yaml: DynVFE -> SPTBackbone -> SSTBEVBackbone -> CenterHead)."""
from .gd_mae import Detector3DTemplate


class CenterPoint(Detector3DTemplate):
    def __init__(self, model_cfg, num_class, dataset, logger=None):
        super().__init__(model_cfg=model_cfg, num_class=num_class, dataset=dataset, logger=logger)
        self.module_topology = ['vfe', 'backbone_3d', 'backbone_2d', 'dense_head']
        self.module_list = self.build_networks()

    def forward(self, batch_dict):
        for cur_module in self.module_list:
            batch_dict = cur_module(batch_dict)
        if self.training:
            loss, tb_dict, disp_dict = self.get_training_loss()
            return {'loss': loss}, tb_dict, disp_dict
        return self.post_processing(batch_dict)

    def get_training_loss(self, sync_for_logging=False):
        """centerpoint.py:24-34; like GDMAE here, tb_dict keeps tensors unless ``sync_for_logging`` (the reference calls .item())"""
        loss_rpn, tb_dict = self.dense_head.get_loss()
        tb_dict = {'loss_rpn': loss_rpn.item() if sync_for_logging else loss_rpn.detach(), **tb_dict}
        return loss_rpn, tb_dict, {}

    def post_processing(self, batch_dict):
        """centerpoint.py:36-50 without the recall bookkeeping (generate_recall_record belongs to the evaluation tooling,
        SURVEY.md section 2: out of scope)"""
        return batch_dict['final_box_dicts'], {}
