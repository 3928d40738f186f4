import sys, ctypes; sys.path.insert(0,'.')
import numpy as np
import bench
from hicexplorer_b200.store import DeviceCSR
from hicexplorer_b200 import _lib
L=_lib.lib()
m_,bins,kw=bench.synth_kwargs('ice100k')
with DeviceCSR.synthetic(bins, **kw) as dev:
    h=dev.handle
    _lib.check(L.hb_dev_ice_begin(h,None))
    st=np.zeros(4,dtype=np.int32); dv=np.zeros(1)
    for k in range(1,400):
        _lib.check(L.hb_dev_ice_spmv(h,None,0,1,None)); _lib.check(L.hb_dev_ice_update(h,None,1,1e-5,None))
        if k%20==0 or k>235 and k<250:
            _lib.check(L.hb_dev_ice_status(h,_lib.ptr(st),None)); _lib.check(L.hb_ice_results(h,None,None,_lib.ptr(dv)))
            print(k, st, dv)
            if st[0]: break
