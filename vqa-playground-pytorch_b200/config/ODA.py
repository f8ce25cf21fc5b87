"""ODA (object-difference attention) — drop-in for the reference's config/ODA.py.

`Model` keeps the constructor, `forward(sample)`, `alpha_dict` and state_dict layout of
config/ODA.py:177-243.  With num_regions = N the attention weight is [4, N*310, 1]
(reference: fuse_dim=11160 = 36*310, config/ODA.py:192).
"""
import os

from ..blocks import MutanFusion, MyATT, MyConv1d, MyLinear, QuestionPassThrough
from ._base import CoreModel

YOUR_DATA_DIR = os.environ.get("VQA_DATA_DIR", "/root/data")
data_dir = os.path.join(YOUR_DATA_DIR, 'VQA/download')
process_dir = os.path.join(YOUR_DATA_DIR, 'VQA/preprocess')
log_dir = os.path.join(YOUR_DATA_DIR, 'VQA/logs')
analyze_dir = os.path.join(YOUR_DATA_DIR, 'VQA/analyze')

version = 2
samplingans = False
loss_metric = "KLD"
vgenome = False
version1_multiple_choices = False
arch = "rcnn"
size = 224

nans = 3000
splitnum = 2
mwc = 0
mql = 26

target_list = ['v', 'q_id', 'q_idxes']
epochs = 70
resume = True
print_freq = 10
lr = 0.0001
load_mem = None
batch_size = 100
clip_grad = True
test_dev_range = None
test_range = None
debug = False

num_regions = 36
precision = "bf16x3"       # fp32-parity arithmetic on bf16 tensor cores (hi + lo operand planes); see _base.CoreModel

method_name = os.path.splitext(os.path.basename(__file__))[0]
if splitnum == 2:
    method_name += '_VAL'
log_dir = os.path.join(log_dir, method_name)
analyze_dir = os.path.join(analyze_dir, method_name)

LAYERS = ["compress_v", "compress_q", "att.conv_att",
          "att.list_linear_v_fusion.0", "att.list_linear_v_fusion.1",
          "att.list_linear_v_fusion.2", "att.list_linear_v_fusion.3",
          "linear_q", "linear_classif"]


class Model(CoreModel):
    MODEL = "ODA"

    def __init__(self, vocab_words=None, num_ans=None, num_regions=num_regions, precision=precision, seq2vec=None):
        super(Model, self).__init__()
        self.vocab_words = vocab_words
        self.num_classes = num_ans
        self.num_regions = num_regions

        self.seq2vec = seq2vec if seq2vec is not None else QuestionPassThrough()
        self.compress_v = MyConv1d(2048, 310, 1, 1, p=0.5, af='relu')
        self.compress_q = MyLinear(2400, 310, p=0.5, af='relu')
        self.att = MyATT(fuse_dim=num_regions * 310, glimpses=4, inputs_dim=2048, att_dim=620, af='relu')
        self.linear_q = MyLinear(2400, 310, p=0.5, af='relu')
        self.fusion_final = MutanFusion(620, 310, 510, 5)
        self.linear_classif = MyLinear(510, self.num_classes, p=0.5)
        self._finish_init(LAYERS, precision)

    def forward(self, sample):
        logits, alpha, _, _ = self._run_core(sample)
        self.alpha_dict = {
            'alphas': alpha[:, :, 0:1]
        }
        return logits
