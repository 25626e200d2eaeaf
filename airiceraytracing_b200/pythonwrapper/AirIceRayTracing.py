"""ctypes loader of libAirIceRayTracing.so, same contract as the reference's pythonwrapper/AirIceRayTracing.py:
the shared object sits next to this file and exports Py_TraceIceToAir(depth, ice, h, d, double[10])."""
import ctypes
import os

dir_path = os.path.dirname(os.path.realpath(__file__))
handle = ctypes.CDLL(os.path.join(dir_path, "libAirIceRayTracing.so"))
handle.Py_TraceIceToAir.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double * 10]
handle.Py_TraceIceToAirBatch.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]


def Py_TraceIceToAir(AntennaDepth, IceLayerHeight, AirTxHeight, HorizontalDistance, ArrayParameters):
    return handle.Py_TraceIceToAir(AntennaDepth, IceLayerHeight, AirTxHeight, HorizontalDistance, ArrayParameters)


def Py_TraceIceToAirBatch(AntennaDepth, IceLayerHeight, AirTxHeight, HorizontalDistance):
    import numpy as np
    h = np.ascontiguousarray(AirTxHeight, dtype=np.float64)
    d = np.ascontiguousarray(HorizontalDistance, dtype=np.float64)
    out = np.empty((h.size, 10))
    rc = handle.Py_TraceIceToAirBatch(AntennaDepth, IceLayerHeight, h.size, h.ctypes.data, d.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("Py_TraceIceToAirBatch failed (%d)" % rc)
    return out
