"""Host-side mirror of the reference's building blocks on the hot path.

Same class names, constructor signatures, parameter names (hence state_dict keys) and error
behaviour as the reference, so checkpoints load unchanged and the parity tests read like the
reference's own smoke blocks:
    MyConv1d, MyLinear, MyATT   config/CoR2.py:56-157 == config/ODA.py:73-174
    Linear, MutanFusion          putils/__init__.py:16-33, :205-241
The nn.Conv1d / nn.Linear children are kept purely as PARAMETER CONTAINERS (same default init,
same keys); their forward is never called — compute goes through ops.py into libvqacore.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID

_ACT = {None: ACT_NONE, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID}


def _act_code(af, who):
    if af not in _ACT:
        raise NotImplementedError("%s: activation %r is not on the CoR2/ODA hot path (relu, sigmoid, None and "
                                  "MyATT's region softmax are)" % (who, af))
    return _ACT[af]


class _DropSite:
    """Dropout call-site bookkeeping shared by MyConv1d / MyLinear: `layer_id` is the Philox
    counter word, assigned by the owning Model in forward-call order (include/vqacore.h)."""
    layer_id = 0
    math = "fp32"
    fixed_seed = None          # tests: pin the Philox key of this site (CoreModel.fixed_seed sets it on every site)

    def _drop_args(self):
        p = float(self.p) if (self.p and self.training) else 0.0
        seed = 0
        if p > 0.0:
            seed = self.fixed_seed if self.fixed_seed is not None else ops.next_seed()
        return p, seed, self.layer_id


class MyConv1d(nn.Module, _DropSite):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, seed=None, p=None, af=None,
                 dim=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.p, self.af, self.dim = p, af, dim
        if kernel_size != 1 or stride != 1 or padding != 0:
            raise NotImplementedError("MyConv1d: only the k=1, stride=1 form used by CoR2/ODA is implemented")
        if seed:
            torch.manual_seed(seed)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding=padding, dilation=1, groups=1,
                              bias=True)

    def forward(self, x):
        if x.dim() != 3:
            raise ValueError('[error] putils.Conv1d(%s, %s, %s, %s): input_dim (%s) should equal to 3' %
                             (self.in_channels, self.out_channels, self.kernel_size, self.stride, x.dim()))
        if self.af == "softmax":
            raise NotImplementedError("MyConv1d(af='softmax') only exists as MyATT.conv_att; call MyATT")
        p, seed, layer = self._drop_args()
        return ops.LinearFn.apply(x, self.conv.weight, self.conv.bias, _act_code(self.af, "MyConv1d"), p, seed, layer,
                                  self.math)


class MyLinear(nn.Module, _DropSite):
    def __init__(self, in_features, out_features, seed=None, p=None, af=None, dim=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.p, self.af, self.dim = p, af, dim
        if seed:
            torch.manual_seed(seed)
        self.linear = nn.Linear(in_features, out_features, bias=True)

    def forward(self, x):
        if x.size()[-1] != self.in_features:
            raise ValueError(
                '[error] putils.Linear(%s, %s): last dimension of input(%s) should equal to in_features(%s)' %
                (self.in_features, self.out_features, x.size(-1), self.in_features))
        p, seed, layer = self._drop_args()
        return ops.LinearFn.apply(x, self.linear.weight, self.linear.bias, _act_code(self.af, "MyLinear"), p, seed,
                                  layer, self.math)


class Linear(nn.Module):
    """putils.Linear: plain linear with the last-dim check (putils/__init__.py:16-33)."""
    math = "fp32"

    def __init__(self, in_features, out_features, bias=True, seed=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        if seed:
            torch.manual_seed(seed)
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        if x.size()[-1] != self.in_features:
            raise ValueError(
                '[error] putils.Linear(%s, %s): last dimension of input(%s) should equal to in_features(%s)' %
                (self.in_features, self.out_features, x.size(-1), self.in_features))
        return ops.LinearFn.apply(x, self.linear.weight, self.linear.bias, ACT_NONE, 0.0, 0, 0, self.math)


class MutanFusion(nn.Module):
    """putils/__init__.py:205-241: sum_r Linear1_r(x1) (.) Linear2_r(x2), per-sample broadcast."""
    math = "fp32"

    def __init__(self, input_dim1, input_dim2, hidden_dim, R, seed=None):
        super().__init__()
        self.input_dim1, self.input_dim2, self.hidden_dim, self.R = input_dim1, input_dim2, hidden_dim, R
        self.list_linear1 = nn.ModuleList([Linear(input_dim1, hidden_dim) for _ in range(R)])
        self.list_linear2 = nn.ModuleList([Linear(input_dim2, hidden_dim) for _ in range(R)])

    def forward(self, inputs1, inputs2):
        if inputs1.size(-1) != self.input_dim1 or inputs2.size(-1) != self.input_dim2:
            raise ValueError('[error] putils.Linear(%s, %s): last dimension of input(%s) should equal to '
                             'in_features(%s)' % (self.input_dim1, self.hidden_dim, inputs1.size(-1), self.input_dim1))
        wb = []
        for lst in (self.list_linear1, self.list_linear2):
            for m in lst:
                wb += [m.linear.weight, m.linear.bias]
        return ops.MutanFn.apply(inputs1, inputs2, self.math, self.R, *wb)


class MyATT(nn.Module):
    """config/CoR2.py:125-157: conv_att -> region softmax -> pooling -> G glimpse linears."""

    def __init__(self, fuse_dim, glimpses, inputs_dim, att_dim, seed=None, af='tanh'):
        super().__init__()
        assert att_dim % glimpses == 0
        if glimpses != 4:
            raise NotImplementedError("MyATT: libvqacore is built for the 4 glimpses CoR2/ODA use")
        self.glimpses, self.inputs_dim, self.att_dim = glimpses, inputs_dim, att_dim
        self.conv_att = MyConv1d(fuse_dim, glimpses, 1, 1, seed=seed, p=0.5, af='softmax', dim=1)
        self.list_linear_v_fusion = nn.ModuleList(
            [MyLinear(inputs_dim, int(att_dim / glimpses), p=0.5, af=af) for _ in range(glimpses)])
        self.af = af

    def forward_with_pooled(self, inputs, fuse):
        """(x_v [B, att_dim], x_att [B,N,G], pooled [B,G,D]); `pooled` (the reference's `tmp`, config/CoR2.py:146) is what
        a following compound-object step needs (its glimpse 0 is sum_i alpha_i x_i)."""
        ca = self.conv_att
        p, seed, layer = ca._drop_args()
        pooled, x_att = ops.RegionSoftmaxPoolFn.apply(inputs, fuse, ca.conv.weight, ca.conv.bias, p, seed, layer)
        list_v = [self.list_linear_v_fusion[g](pooled[:, g, :]) for g in range(self.glimpses)]
        return torch.cat(list_v, 1), x_att, pooled

    def forward(self, inputs, fuse):
        x_v, x_att, _ = self.forward_with_pooled(inputs, fuse)
        return x_v, torch.split(x_att, 1, dim=2)


class QuestionPassThrough(nn.Module):
    """Stand-in for the question encoder: `sample['q_idxes']` already holds the 2400-d embedding.
    The reference's SkipThoughts/BayesianGRU (putils/__init__.py:878-985) is outside the hot path
    (BASELINE.json treats ques_emb as an input); pass any nn.Module producing
    [B,2400] as `seq2vec=` to restore it."""

    def forward(self, q):
        return q


# ----------------------------------------------------------------------------------------- question encoder
class BayesianGRUCell(nn.Module):
    """Parameter container of putils.BayesianGRUCell (putils/__init__.py:604-637 over AbstractGRUCell :566-583): six
    nn.Linear children with the reference's names (weight_ir/ii/in with bias, weight_hr/hi/hn without), so the
    state_dict keys `gru_cell.weight_*.{weight,bias}` that SkipThoughts.load_bayesiangru_state_dict fills are the same."""

    def __init__(self, input_size, hidden_size, bias_ih=True, bias_hh=False, dropout=0.25, af='tanh'):
        super().__init__()
        if bias_hh or not bias_ih:
            raise NotImplementedError("BayesianGRUCell: the SkipThoughts form (bias_ih=True, bias_hh=False) is implemented")
        self.input_size, self.hidden_size, self.dropout, self.af = input_size, hidden_size, dropout, af
        self.weight_ir = nn.Linear(input_size, hidden_size, bias=True)
        self.weight_ii = nn.Linear(input_size, hidden_size, bias=True)
        self.weight_in = nn.Linear(input_size, hidden_size, bias=True)
        self.weight_hr = nn.Linear(hidden_size, hidden_size, bias=False)
        self.weight_hi = nn.Linear(hidden_size, hidden_size, bias=False)
        self.weight_hn = nn.Linear(hidden_size, hidden_size, bias=False)

    def set_dropout(self, dropout):
        self.dropout = dropout


class BayesianGRU(nn.Module):
    """putils.BayesianGRU (putils/__init__.py:660-741) with return_last=True: the hidden state at each sequence's last
    non-PAD position.  `forward(emb_or_idx, ...)` is not exposed step by step: SkipThoughts below runs the whole
    encoder as one autograd node (ops.BayesianGruFn)."""

    def __init__(self, input_size, hidden_size, bias_ih=True, bias_hh=False, dropout=0.25, return_last=True, af='tanh'):
        super().__init__()
        if return_last is not True:
            raise NotImplementedError("BayesianGRU: return_last=True (the form config/CoR2.py and config/ODA.py use)")
        self.input_size, self.hidden_size = input_size, hidden_size
        self.dropout, self.return_last, self.af = dropout, return_last, af
        self.gru_cell = BayesianGRUCell(input_size, hidden_size, bias_ih, bias_hh, dropout=dropout, af=af)

    def set_dropout(self, dropout):
        self.dropout = dropout
        self.gru_cell.set_dropout(dropout)


class SkipThoughts(nn.Module):
    """putils.SkipThoughts (putils/__init__.py:878-985) without the downloads: nn.Embedding(len(vocab), 620,
    padding_idx=0) + BayesianGRU(620, 2400, dropout=0.25); state_dict keys `embedding.weight`,
    `gru.gru_cell.weight_*` as in the reference, so a reference checkpoint's `seq2vec.*` entries load unchanged.
    The pretrained uni-skip tables are not shipped (no network): weights keep their default init unless loaded.
    forward(q_idxes [B, T] int64, 0 = PAD) -> [B, 2400]."""

    def __init__(self, vocab_list, data_dir=None, gru='BayesianGRU', return_last=True, af='tanh', precision="bf16x3"):
        super().__init__()
        if gru != 'BayesianGRU':
            raise ValueError
        self.vocab_list, self.data_dir, self.af, self.math = vocab_list, data_dir, af, precision
        self.embedding = nn.Embedding(num_embeddings=len(vocab_list), embedding_dim=620, padding_idx=0)
        self.gru = BayesianGRU(input_size=620, hidden_size=2400, dropout=0.25, return_last=return_last, af=af)
        self.fixed_seed = None
        self.graph_capture = False      # set by the owning CoreModel while engine.GraphedStep captures / replays it

    def forward(self, x, return_hidden=False):
        c = self.gru.gru_cell
        p = float(self.gru.dropout) if self.training else 0.0
        seed = 0
        if p > 0.0:
            if self.graph_capture:
                # the sequence-tied masks are drawn from a host-side key: a captured step would replay ONE mask forever
                raise NotImplementedError("SkipThoughts: train-mode dropout inside a CUDA-graph-captured step is not "
                                          "supported (the encoder's Philox key is host-resident); run the encoder "
                                          "eagerly in front of the graphed core, or in eval mode")
            seed = self.fixed_seed if self.fixed_seed is not None else ops.next_seed()
        out = ops.BayesianGruFn.apply(x, self.embedding.weight, c.weight_ir.weight, c.weight_ir.bias, c.weight_ii.weight,
                                      c.weight_ii.bias, c.weight_in.weight, c.weight_in.bias, c.weight_hr.weight,
                                      c.weight_hi.weight, c.weight_hn.weight, p, seed, self.af, self.math)
        return out
