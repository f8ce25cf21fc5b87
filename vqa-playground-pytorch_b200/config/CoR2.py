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


def layers_for(steps):
    """Dropout call sites in forward-call order for a chain of `steps` attention steps (steps = 2: LAYERS, the stock
    model).  Every further step s adds its two gates, its compress_v{s} and its att{s}, before linear_q."""
    out = LAYERS[:17]
    for s in range(3, steps + 1):
        out += ["compress_q_%d" % (2 * s - 3), "expand_q_%d" % (2 * s - 3), "compress_q_%d" % (2 * s - 2),
                "expand_q_%d" % (2 * s - 2), "compress_v%d" % s, "att%d.conv_att" % s]
        out += ["att%d.list_linear_v_fusion.%d" % (s, g) for g in range(4)]
    return out + LAYERS[17:]


class Model(CoreModel):
    """config/CoR2.py:160-240.  `steps` = number of attention steps of the chain of reasoning: 2 is the shipped model
    (att1 -> compound objects -> att2) and runs as ONE fused plan in libvqacore; steps >= 3 repeats the relational step
    (the 3-step model behind the `alpha3` / `v3_feature` keys of visu.py:2491-2495 and
    CoR_Visulization.py:107-109, whose config the reference does not ship): step s builds its compound objects from the
    previous step's attention (block1 = the previous objects, block2 = v, as decare_cat(v, v, q) does for step 2),
    compresses them, fuses with the question and attends.  Those models are composed from the same building blocks
    (blocks.py -> ops.py -> libvqacore kernels) in Python; their parity oracle is the reference's own decare_cat / MyATT
    composed once more (oracle.reasoning_core.cor_forward), labelled UNPINNED (no reference model exists to pin it)."""
    MODEL = "CoR2"

    def __init__(self, vocab_words=None, num_ans=None, num_regions=num_regions, precision=precision, seq2vec=None,
                 steps=2, compose=False):
        super(Model, self).__init__()
        if steps < 2:
            raise ValueError("CoR needs at least the two attention steps of config/CoR2.py")
        self.vocab_words = vocab_words
        self.num_classes = num_ans
        self.num_regions = num_regions
        self.steps = steps
        self.compose = bool(compose) or steps != 2      # compose=True: the Python-composed chain for steps = 2 as well (tests)

        self.seq2vec = seq2vec if seq2vec is not None else QuestionPassThrough()
        self.compress_v = MyConv1d(2048, 310, 1, 1, p=0.5, af='relu')
        self.compress_v2 = MyConv1d(2048, 310, 1, 1, p=0.5, af='relu')
        self.compress_q = MyLinear(2400, 310, p=0.5, af='relu')

        self.fusion_vq1 = MutanFusion(310, 310, 510, 2)
        self.att1 = MyATT(fuse_dim=510, glimpses=4, inputs_dim=2048, att_dim=620, af='relu')

        self.fusion_vq2 = MutanFusion(310, 310, 510, 2)
        self.att2 = MyATT(fuse_dim=510, glimpses=4, inputs_dim=2048, att_dim=620, af='relu')

        self.linear_q = MyLinear(2400, 310, p=0.5, af='relu')
        self.fusion_final = MutanFusion(620 * steps, 310, 510, 2)
        self.linear_classif = MyLinear(510, self.num_classes, p=0.5)

        self.compress_q_1 = MyLinear(2400, 310, p=0.5, af='relu')
        self.expand_q_1 = MyLinear(310, 2048, p=0.5, af='sigmoid')

        self.compress_q_2 = MyLinear(2400, 310, p=0.5, af='relu')
        self.expand_q_2 = MyLinear(310, 2048, p=0.5, af='sigmoid')
        for s in range(3, steps + 1):                   # further steps: registered after the stock parameters
            setattr(self, "compress_v%d" % s, MyConv1d(2048, 310, 1, 1, p=0.5, af='relu'))
            setattr(self, "fusion_vq%d" % s, MutanFusion(310, 310, 510, 2))
            setattr(self, "att%d" % s, MyATT(fuse_dim=510, glimpses=4, inputs_dim=2048, att_dim=620, af='relu'))
            for k in (2 * s - 3, 2 * s - 2):
                setattr(self, "compress_q_%d" % k, MyLinear(2400, 310, p=0.5, af='relu'))
                setattr(self, "expand_q_%d" % k, MyLinear(310, 2048, p=0.5, af='sigmoid'))
        self._finish_init(layers_for(steps), precision)

    def _forward_composed(self, sample):
        """The chain of config/CoR2.py:201-237 for any number of steps, block by block."""
        from .. import ops
        N = self.num_regions
        v = sample['v'].contiguous().view(-1, N, 2048)
        q = self.seq2vec(sample['q_idxes'])
        for m in self.modules():
            if hasattr(m, "layer_id"):
                m.fixed_seed = self.fixed_seed
        ql = self.compress_q(q)
        x, feats, alphas, objects = v, [], [], []
        pooled = alpha = None
        for s in range(1, self.steps + 1):
            if s > 1:      # compound objects from the previous step's attention (decare_cat + the alpha[0]-weighted sum)
                g1 = getattr(self, "expand_q_%d" % (2 * s - 3))(getattr(self, "compress_q_%d" % (2 * s - 3))(q))
                g2 = getattr(self, "expand_q_%d" % (2 * s - 2))(getattr(self, "compress_q_%d" % (2 * s - 2))(q))
                x = ops.CorCompoundFn.apply(v, pooled, alpha, g1, g2)
                objects.append(x)
            xl = getattr(self, "compress_v" if s == 1 else "compress_v%d" % s)(x)
            fuse = getattr(self, "fusion_vq%d" % s)(xl, ql)
            x_att, alpha, pooled = getattr(self, "att%d" % s).forward_with_pooled(x, fuse)
            feats.append(x_att)
            alphas.append(alpha)
        v_f = torch.cat(feats, dim=1)
        x = self.fusion_final(v_f, self.linear_q(q))
        logits = self.linear_classif(x)
        self.alpha_dict = {'alpha%d' % (s + 1): torch.split(a.detach(), 1, dim=2) for s, a in enumerate(alphas)}
        self.alpha_dict['feature'] = objects[0].detach()[:, 0:2, :]
        for s, o in enumerate(objects[1:], start=3):
            self.alpha_dict['v%d_feature' % s] = o.detach()[:, 0:2, :]
        return logits

    def forward(self, sample):
        if self.compose:
            return self._forward_composed(sample)
        logits, alpha1, alpha2, v2 = self._run_core(sample)
        self.alpha_dict = {
            'alpha1': torch.split(alpha1, 1, dim=2),
            'alpha2': torch.split(alpha2, 1, dim=2),
            'feature': v2[:, 0:2, :]          # reference: v2_feature[:, [0, 1], :] (config/CoR2.py:224); a view, no kernel
        }
        return logits
