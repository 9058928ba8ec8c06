"""iodine_b200 -- B200-native engine for IODINE's iterative-refinement inference loop.

Host side mirrors the reference's Python surface (``lib/modeling``: ``make_model``,
``IODINE``); the loop itself is hand-written sm_100a CUDA behind a C ABI
(``include/iodine_b200.h``, ``iodine_b200/lib/libiodine_b200.so``).
"""
__version__ = '0.1.0'
