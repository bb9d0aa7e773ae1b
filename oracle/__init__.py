"""CPU oracle for the U-Net hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (oct_image_segmentation_models_b200)
never does, and fails loudly when its CUDA library is missing.
"""
