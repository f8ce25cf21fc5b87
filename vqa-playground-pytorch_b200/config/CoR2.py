"""CoR2 (chain of reasoning) — drop-in for the reference's config/CoR2.py.

Module-level names are the ones train.py reads (train.py:323-447, :487-517); `Model` keeps the
constructor, `forward(sample)`, `alpha_dict` and state_dict layout of config/CoR2.py:160-240.
"""
import os

import torch

from ..blocks import MutanFusion, MyATT, MyConv1d, MyLinear, QuestionPassThrough
from ._base import CoreModel

YOUR_DATA_DIR = os.environ.get("VQA_DATA_DIR", "/mnt/cephfs/lab/liujinlai.licio")
data_dir = os.path.join(YOUR_DATA_DIR, 'data/VQA/download')
process_dir = os.path.join(YOUR_DATA_DIR, 'data/VQA/preprocess')
log_dir = os.path.join(YOUR_DATA_DIR, 'data/VQA/logs')
analyze_dir = os.path.join(YOUR_DATA_DIR, 'data/VQA/analyze')

version = 2
samplingans = False
loss_metric = "KLD"
vgenome = False
version1_multiple_choices = False
arch = "rcnn"
size = 224

nans = 2000
splitnum = 2
mwc = 0
mql = 26

target_list = ['v', 'q_id', 'q_idxes']
epochs = 70
restart_epoch = None
keeping_epoch = 40

resume = True
print_freq = 10
lr = 0.0001
load_mem = None
batch_size = 100
clip_grad = True
test_dev_range = None
test_range = None
debug = False

# additions (reference defaults)
num_regions = 36
precision = "bf16x3"       # fp32-parity arithmetic on bf16 tensor cores (hi + lo operand planes); see _base.CoreModel

method_name = os.path.splitext(os.path.basename(__file__))[0]
if splitnum == 2:
    method_name += '_VAL'
log_dir = os.path.join(log_dir, method_name)
analyze_dir = os.path.join(analyze_dir, method_name)

LAYERS = ["compress_q", "compress_v", "att1.conv_att",
          "att1.list_linear_v_fusion.0", "att1.list_linear_v_fusion.1",
          "att1.list_linear_v_fusion.2", "att1.list_linear_v_fusion.3",
          "compress_q_1", "expand_q_1", "compress_q_2", "expand_q_2",
          "compress_v2", "att2.conv_att",
          "att2.list_linear_v_fusion.0", "att2.list_linear_v_fusion.1",
          "att2.list_linear_v_fusion.2", "att2.list_linear_v_fusion.3",
          "linear_q", "linear_classif"]


class Model(CoreModel):
    MODEL = "CoR2"

    def __init__(self, vocab_words=None, num_ans=None, num_regions=num_regions, precision=precision, seq2vec=None):
        super(Model, self).__init__()
        self.vocab_words = vocab_words
        self.num_classes = num_ans
        self.num_regions = num_regions

        self.seq2vec = seq2vec if seq2vec is not None else QuestionPassThrough()
        self.compress_v = MyConv1d(2048, 310, 1, 1, p=0.5, af='relu')
        self.compress_v2 = MyConv1d(2048, 310, 1, 1, p=0.5, af='relu')
        self.compress_q = MyLinear(2400, 310, p=0.5, af='relu')

        self.fusion_vq1 = MutanFusion(310, 310, 510, 2)
        self.att1 = MyATT(fuse_dim=510, glimpses=4, inputs_dim=2048, att_dim=620, af='relu')

        self.fusion_vq2 = MutanFusion(310, 310, 510, 2)
        self.att2 = MyATT(fuse_dim=510, glimpses=4, inputs_dim=2048, att_dim=620, af='relu')

        self.linear_q = MyLinear(2400, 310, p=0.5, af='relu')
        self.fusion_final = MutanFusion(1240, 310, 510, 2)
        self.linear_classif = MyLinear(510, self.num_classes, p=0.5)

        self.compress_q_1 = MyLinear(2400, 310, p=0.5, af='relu')
        self.expand_q_1 = MyLinear(310, 2048, p=0.5, af='sigmoid')

        self.compress_q_2 = MyLinear(2400, 310, p=0.5, af='relu')
        self.expand_q_2 = MyLinear(310, 2048, p=0.5, af='sigmoid')
        self._finish_init(LAYERS, precision)

    def forward(self, sample):
        logits, alpha1, alpha2, v2 = self._run_core(sample)
        self.alpha_dict = {
            'alpha1': torch.split(alpha1, 1, dim=2),
            'alpha2': torch.split(alpha2, 1, dim=2),
            'feature': v2[:, 0:2, :]          # reference: v2_feature[:, [0, 1], :] (config/CoR2.py:224); a view, no kernel
        }
        return logits
