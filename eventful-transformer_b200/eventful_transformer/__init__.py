"""
eventful_transformer -- B200-native drop-in for the reference package of the same name.

Same import paths, class names, constructor kwargs, state-dict keys and control
API (reset / counting / set-policy by attribute) as WISION-Lab/eventful-transformer's
`eventful_transformer` package, so `models/vitdet.py` and `models/vivit.py` run on top of
it unchanged; the arithmetic underneath is libeventful_b200.so (hand-written sm_100a
kernels behind the C ABI in include/eventful_b200.h).  CUDA only -- no CPU fallback.
"""
