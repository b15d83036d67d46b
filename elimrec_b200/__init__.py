"""elimrec_b200 - B200-native training + full-ranking hot path of EliMRec.

Host side mirrors the reference's Python interface (models.EliMRec, evaluator.ProxyEvaluator,
data.PairwiseSamplerV2); all arithmetic runs in hand-written sm_100a kernels behind the C-ABI in
``include/elimrec_b200.h`` (``elimrec_b200/csrc/libelimrec_b200.so``).  No CPU fallback.
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not need torch.cuda or the .so
    if name in ("EliMRec", "BasicModel"):
        from . import model
        return getattr(model, name)
    if name in ("ProxyEvaluator", "UniEvaluator"):
        from . import evaluator
        return getattr(evaluator, name)
    if name in ("PairwiseSamplerV2",):
        from . import sampler
        return getattr(sampler, name)
    if name in ("Dataset", "Config"):
        from . import data
        return getattr(data, name)
    raise AttributeError(name)
