"""upside-md_b200: B200-native implementation of Upside's MD inner loop behind the reference's own interfaces.

Modules
  upside_engine  ctypes binding to libupside_b200.so - same names as the reference's py/upside_engine.py
                 (`Upside`, `in_process_upside`, clamped-spline helpers) plus `BatchEngine`, the batched engine
  config         Python-3 restatement of the ff_1 part of py/upside_config.py
  h5lite         minimal HDF5 reader/writer (no libhdf5/h5py in the image)
"""
