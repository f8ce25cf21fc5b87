"""B200 (sm_100a) reasoning core of bupt-cist/vqa-playground-pytorch's CoR2 and ODA models.

    from vqa_playground_pytorch_b200.config import CoR2
    model = CoR2.Model(vocab_words, num_ans).cuda()      # same constructor / forward(sample) / state_dict
    logits = model({'v': v, 'q_idxes': q_emb})

Compute lives in libvqacore_sm100a.so (csrc/, C ABI in include/vqacore.h); this package is the
host-side mirror of the reference's Python interface for that path.  No CPU fallback exists.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "blocks", "config", "parallel", "engine", "optim"]
