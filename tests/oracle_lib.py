"""ctypes binding of oracle/liboracle.so (CPU oracle; TEST INFRASTRUCTURE, never a product path)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "liboracle.so")


class CfBasis(C.Structure):
    _fields_ = [("nshell", C.c_int), ("type", C.POINTER(C.c_int)), ("nprim", C.POINTER(C.c_int)),
                ("prim_offset", C.POINTER(C.c_int)), ("exps", C.POINTER(C.c_double)),
                ("coefs_normalized", C.POINTER(C.c_double)), ("center_xyz", C.POINTER(C.c_double)),
                ("shell2atom", C.POINTER(C.c_int))]


def as_cf_basis(fb):
    """FlatBasis -> (CfBasis, keepalive)"""
    arrs = dict(type=np.ascontiguousarray(fb.type, np.int32), nprim=np.ascontiguousarray(fb.nprim, np.int32),
                prim_offset=np.ascontiguousarray(fb.prim_offset, np.int32),
                exps=np.ascontiguousarray(fb.exps, np.float64),
                coefs=np.ascontiguousarray(fb.coefs_normalized, np.float64),
                xyz=np.ascontiguousarray(fb.center_xyz, np.float64).reshape(-1),
                s2a=np.ascontiguousarray(fb.shell2atom, np.int32))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    b = CfBasis(fb.nshell, ip(arrs["type"]), ip(arrs["nprim"]), ip(arrs["prim_offset"]), dp(arrs["exps"]),
                dp(arrs["coefs"]), dp(arrs["xyz"]), ip(arrs["s2a"]))
    return b, arrs


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _fmat(a):
    return None if a is None else np.asfortranarray(a, dtype=np.float64)


class Oracle:
    def __init__(self):
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        self.lib = C.CDLL(_SO)
        L = self.lib
        L.oracle_nbf.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        L.ref_store_build.restype = C.c_void_p
        L.ref_store_len.restype = C.c_long
        L.ref_store_len.argtypes = [C.c_void_p]
        L.ref_store_free.argtypes = [C.c_void_p]
        L.oracle_set_threads.restype = C.c_int
        self.nthreads = L.oracle_max_threads()

    def set_threads(self, n=0):
        """OpenMP threads of every later call; n <= 0 = all online cores, whatever OMP_NUM_THREADS a launcher exported
        (torchrun sets it to 1)."""
        self.nthreads = int(self.lib.oracle_set_threads(C.c_int(n)))
        return self.nthreads

    def boys(self, mmax, T):
        F = np.zeros(mmax + 1)
        self.lib.oracle_boys(C.c_int(mmax), C.c_double(T), _dp(F))
        return F

    def pure_matrix(self, l):
        Cm = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2))
        self.lib.oracle_pure_matrix(C.c_int(l), _dp(Cm))
        return Cm

    def eri_quartet(self, fb, s1, s2, s3, s4):
        b, keep = as_cf_basis(fb)
        n = [int(fb.nfun[s]) for s in (s1, s2, s3, s4)]
        buf = np.zeros(n)
        self.lib.oracle_eri_shell_quartet(C.byref(b), s1, s2, s3, s4, _dp(buf))
        return buf

    def eri_full(self, fb):
        """dense (ij|kl) tensor, small systems only"""
        nbf = fb.nbf
        out = np.zeros((nbf,) * 4)
        o = fb.shell2bf
        n = fb.nfun
        for a in range(fb.nshell):
            for b_ in range(a + 1):
                for c in range(a + 1):
                    for d in range(c + 1):
                        blk = self.eri_quartet(fb, a, b_, c, d)
                        sa, sb, sc, sd = (slice(o[s], o[s] + n[s]) for s in (a, b_, c, d))
                        out[sa, sb, sc, sd] = blk
                        out[sb, sa, sc, sd] = blk.transpose(1, 0, 2, 3)
                        out[sa, sb, sd, sc] = blk.transpose(0, 1, 3, 2)
                        out[sb, sa, sd, sc] = blk.transpose(1, 0, 3, 2)
                        out[sc, sd, sa, sb] = blk.transpose(2, 3, 0, 1)
                        out[sd, sc, sa, sb] = blk.transpose(3, 2, 0, 1)
                        out[sc, sd, sb, sa] = blk.transpose(2, 3, 1, 0)
                        out[sd, sc, sb, sa] = blk.transpose(3, 2, 1, 0)
        return out

    def one_electron(self, fb, Z, xyz_bohr):
        b, keep = as_cf_basis(fb)
        nbf = fb.nbf
        S, T, V = (np.zeros((nbf, nbf), order="F") for _ in range(3))
        Zd = np.ascontiguousarray(Z, np.float64)
        R = np.ascontiguousarray(xyz_bohr, np.float64)
        self.lib.oracle_one_electron(C.byref(b), C.c_int(len(Zd)), _dp(Zd), _dp(R), _dp(S), _dp(T), _dp(V))
        return S, T, V

    def repulsion_diag(self, fb):
        b, keep = as_cf_basis(fb)
        d = np.zeros((fb.nbf, fb.nbf), order="F")
        self.lib.ref_getRepulsionDiag(C.byref(b), _dp(d))
        return d

    def _outs(self, nbf, Dd, Da, Db):
        J = np.zeros((nbf, nbf), order="F")
        Ks = [np.zeros((nbf, nbf), order="F") if D is not None else None for D in (Dd, Da, Db)]
        return J, Ks

    def reference_jk(self, fb, Dd=None, Da=None, Db=None, exx=1.0, threshold=-1.0, nthreads=None):
        """The reference's stored-integral path, restated literally. -> J, Kd, Ka, Kb, (RepulsionLength, ShellQuartetLength)"""
        b, keep = as_cf_basis(fb)
        Dd, Da, Db = _fmat(Dd), _fmat(Da), _fmat(Db)
        J, Ks = self._outs(fb.nbf, Dd, Da, Db)
        counts = (C.c_long * 2)()
        self.lib.ref_full_path(C.byref(b), C.c_double(threshold), C.c_double(exx), C.c_int(nthreads or self.nthreads),
                               _dp(Dd), _dp(Da), _dp(Db), _dp(J), _dp(Ks[0]), _dp(Ks[1]), _dp(Ks[2]), counts)
        return J, Ks[0], Ks[1], Ks[2], (counts[0], counts[1])

    def direct_jk(self, fb, Dd=None, Da=None, Db=None, exx=1.0, nthreads=None, stride=1, offset=0):
        b, keep = as_cf_basis(fb)
        Dd, Da, Db = _fmat(Dd), _fmat(Da), _fmat(Db)
        J, Ks = self._outs(fb.nbf, Dd, Da, Db)
        counts = (C.c_long * 2)()
        self.lib.oracle_direct_jk(C.byref(b), C.c_int(fb.nbf), _dp(Dd), _dp(Da), _dp(Db), C.c_double(exx),
                                  _dp(J), _dp(Ks[0]), _dp(Ks[1]), _dp(Ks[2]), C.c_int(nthreads or self.nthreads),
                                  C.c_int(stride), C.c_int(offset), counts)
        return J, Ks[0], Ks[1], Ks[2], (counts[0], counts[1])

    def jk_block(self, fb, Dtot, D, sa, sb, nthreads=None):
        b, keep = as_cf_basis(fb)
        Dtot, D = _fmat(Dtot), _fmat(D)
        na, nb = int(fb.nfun[sa]), int(fb.nfun[sb])
        Jb, Kb = np.zeros((na, nb)), np.zeros((na, nb))
        self.lib.oracle_jk_block(C.byref(b), C.c_int(fb.nbf), _dp(Dtot), _dp(D), C.c_int(sa), C.c_int(sb), _dp(Jb), _dp(Kb),
                                 C.c_int(nthreads or self.nthreads))
        return Jb, Kb

    # first-derivative path
    def eri_deriv_quartet(self, fb, s1, s2, s3, s4):
        """[12 = centre of s1..s4 x (x,y,z)][n1][n2][n3][n4]: the 12 buffers libint2 hands the reference (Int4C2E.cpp:377-389)"""
        b, keep = as_cf_basis(fb)
        n = [int(fb.nfun[s]) for s in (s1, s2, s3, s4)]
        buf = np.zeros([12] + n)
        self.lib.oracle_eri_deriv_quartet(C.byref(b), s1, s2, s3, s4, _dp(buf))
        return buf

    def grad_matrices(self, fb, D, exx=1.0, nthreads=None):
        """getRepulsion1 (Int4C2E.cpp:312-408): 3*natom matrices G^(atom,xyz)[D], each nbf x nbf"""
        b, keep = as_cf_basis(fb)
        natom = int(np.max(fb.shell2atom)) + 1
        D = _fmat(D)
        G = np.zeros((3 * natom, fb.nbf * fb.nbf))
        self.lib.ref_getRepulsion1(C.byref(b), C.c_int(natom), _dp(D), C.c_double(exx), _dp(G), C.c_int(nthreads or self.nthreads))
        return [np.asfortranarray(g.reshape(fb.nbf, fb.nbf).T) for g in G]

    def contract_grads(self, fb, D1, D2, exx=1.0, nthreads=None):
        """Int4C2E::ContractGrads(D1, D2) (Int4C2E.cpp:747-763) -> [3*natom]"""
        b, keep = as_cf_basis(fb)
        natom = int(np.max(fb.shell2atom)) + 1
        D1, D2 = _fmat(D1), _fmat(D2)
        g = np.zeros(3 * natom)
        self.lib.ref_ContractGrads(C.byref(b), C.c_int(natom), _dp(D1), _dp(D2), C.c_double(exx), _dp(g),
                                   C.c_int(nthreads or self.nthreads))
        return g

    # second-derivative path (tail)
    def eri_deriv2_quartet(self, fb, s1, s2, s3, s4):
        """[78 = upper triangle of the 12 x 12 (centre, xyz) matrix, row-major][n1][n2][n3][n4]: the buffers libint2 hands the
        reference for deriv_order 2 (Int4C2E.cpp:468-472)"""
        b, keep = as_cf_basis(fb)
        n = [int(fb.nfun[s]) for s in (s1, s2, s3, s4)]
        buf = np.zeros([78] + n)
        self.lib.oracle_eri_deriv2_quartet(C.byref(b), s1, s2, s3, s4, _dp(buf))
        return buf

    def contract_hess(self, fb, D, exx=1.0, nthreads=None):
        """getRepulsion2 / Int4C2E::ContractHesss (Int4C2E.cpp:410-492, :792-811) -> [3*natom][3*natom]"""
        b, keep = as_cf_basis(fb)
        natom = int(np.max(fb.shell2atom)) + 1
        D = _fmat(D)
        H = np.zeros((3 * natom, 3 * natom), order="F")
        self.lib.ref_getRepulsion2(C.byref(b), C.c_int(natom), _dp(D), C.c_double(exx), _dp(H), C.c_int(nthreads or self.nthreads))
        return H

    # stored-integral handle (B1 timing)
    def store_build(self, fb, threshold=-1.0):
        b, keep = as_cf_basis(fb)
        h = self.lib.ref_store_build(C.byref(b), C.c_double(threshold))
        return h

    def store_len(self, h):
        return self.lib.ref_store_len(h)

    def store_contract(self, h, nbf, Dd=None, Da=None, Db=None, exx=1.0, nthreads=None):
        Dd, Da, Db = _fmat(Dd), _fmat(Da), _fmat(Db)
        J, Ks = self._outs(nbf, Dd, Da, Db)
        self.lib.ref_store_contract(C.c_void_p(h), _dp(Dd), _dp(Da), _dp(Db), C.c_double(exx), C.c_int(nthreads or self.nthreads),
                                    _dp(J), _dp(Ks[0]), _dp(Ks[1]), _dp(Ks[2]))
        return J, Ks[0], Ks[1], Ks[2]

    def store_free(self, h):
        self.lib.ref_store_free(C.c_void_p(h))
