"""fib_tf_b200 -- B200-native (sm_100a) explicit time-stepper for 2-D cardiac monodomain models,
a drop-in for the hot path of siravan/fib_tf behind the reference's own IonicModel API.
The compute lives in libfibb200.so (hand-written CUDA behind the C ABI of include/fib_b200.h);
this package is the thin Python host side.  No TensorFlow, no Triton, no CPU fallback."""
__version__ = '0.1.0'
