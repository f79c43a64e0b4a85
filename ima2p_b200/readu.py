"""Host mirror of the `.u` reader of the C ABI (ima2p_dataset_*): readdata, readata.cpp:1038-1123."""
import ctypes as C

import numpy as np

from . import capi


def read_u(path, lib=None):
    """Parse a reference input file; returns dict(npops, tree, loci=[dict(model, numgenes, numsites, totsites, numbases,
    nlinked, hval, samppop, name, seq[numgenes][numsites], mult, A[nlinked][numgenes], minA, maxA, pi, urate)])."""
    l = lib or capi.lib()
    h = C.c_void_p()
    capi.check(l, l.ima2p_dataset_read(str(path).encode(), C.byref(h)))
    try:
        npops, nloci = C.c_int(), C.c_int()
        tree = C.create_string_buffer(256)
        capi.check(l, l.ima2p_dataset_dims(h, C.byref(npops), C.byref(nloci), tree, 256))
        loci = []
        for li in range(nloci.value):
            info = (C.c_int * 8)()
            hval = C.c_double()
            samp = (C.c_int * npops.value)()
            name = C.create_string_buffer(64)
            capi.check(l, l.ima2p_dataset_locus(h, li, info, C.byref(hval), samp, name, 64))
            model, n, ns, tot, nb, nlinked, nur = info[0], info[1], info[2], info[3], info[4], info[5], info[6]
            seq, mult = np.zeros((n, ns), np.int32), np.zeros(ns, np.int32)
            A, minA, maxA = np.zeros((nlinked, n), np.int32), np.zeros(nlinked, np.int32), np.zeros(nlinked, np.int32)
            pi, ur = np.zeros(4), np.zeros(max(nur, 1))
            ip = lambda a: a.ctypes.data_as(capi.c_int_p)
            capi.check(l, l.ima2p_dataset_locus_data(h, li, ip(seq), ip(mult), ip(A), ip(minA), ip(maxA),
                                                     pi.ctypes.data_as(capi.c_dbl_p), ur.ctypes.data_as(capi.c_dbl_p)))
            loci.append(dict(model=model, numgenes=n, numsites=ns, totsites=tot, numbases=nb, nlinked=nlinked, hval=hval.value,
                             samppop=list(samp), name=name.value.decode(), seq=seq, mult=mult, A=A, minA=minA, maxA=maxA, pi=pi,
                             urate=ur[:nur]))
        return dict(npops=npops.value, tree=tree.value.decode(), loci=loci)
    finally:
        l.ima2p_dataset_free(h)


def ti_create(path, header="", lib=None):
    l = lib or capi.lib()
    capi.check(l, l.ima2p_ti_create(str(path).encode(), header.encode()))


def ti_append(path, rows, lib=None):
    l = lib or capi.lib()
    r = np.ascontiguousarray(rows, dtype=np.float32)
    capi.check(l, l.ima2p_ti_append(str(path).encode(), r.ctypes.data_as(capi.c_flt_p), r.shape[0], r.shape[1]))


def ti_load(path, rowlen, max_rows=None, lib=None):
    """Rows of a .ti file as float32 [G][rowlen] (loadgenealogyvalues)."""
    l = lib or capi.lib()
    n = C.c_longlong()
    capi.check(l, l.ima2p_ti_load(str(path).encode(), rowlen, None, 0, C.byref(n)))
    want = n.value if max_rows is None else min(n.value, max_rows)
    rows = np.zeros((want, rowlen), np.float32)
    capi.check(l, l.ima2p_ti_load(str(path).encode(), rowlen, rows.ctypes.data_as(capi.c_flt_p), want, C.byref(n)))
    return rows[:n.value]
