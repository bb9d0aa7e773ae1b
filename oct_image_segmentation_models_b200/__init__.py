"""B200-native drop-in for the dense compute path of oct_image_segmentation_models.

Public surface mirrors the reference package (models / prediction / evaluation /
training / common); the arithmetic runs in csrc/liboctseg.so (hand-written sm_100a
CUDA) through ctypes.  There is no TensorFlow and no CPU fallback.
"""
__version__ = "0.1.0"
