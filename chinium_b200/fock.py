"""Python mirror of the reference's `class Int4C2E` (src/Integral/Int4C2E.h:10-50) over the C ABI.

Same method names, argument meaning and error behaviour as the reference:
    Int4C2E(basis, exx, threshold); getRepulsionDiag(output); getRepulsionLength(output);
    getRepulsionIndices(output); getThreadPointers(nthreads, output); CalculateIntegrals(order, output);
    ContractInts(Dd, Da, Db, nthreads, output) -> (J, Kd, Ka, Kb)     [None = the reference's 0x0 matrix]
    ContractInts([D...], nthreads, output) -> [G...]
    ContractGrads(D1, D2, output) -> [3*natoms]
Everything numerical happens in libchinium_fock.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

# CHINIUM_FOCK_LIB: developer override used for A/B measurements of kernel variants (still an in-tree CUDA build)
LIB_PATH = os.environ.get("CHINIUM_FOCK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libchinium_fock.so")


class FockEngineError(RuntimeError):
    pass


class _CfBasis(C.Structure):
    _fields_ = [("nshell", C.c_int), ("type", C.POINTER(C.c_int)), ("nprim", C.POINTER(C.c_int)),
                ("prim_offset", C.POINTER(C.c_int)), ("exps", C.POINTER(C.c_double)),
                ("coefs_normalized", C.POINTER(C.c_double)), ("center_xyz", C.POINTER(C.c_double)),
                ("shell2atom", C.POINTER(C.c_int))]


class _CfOptions(C.Structure):
    _fields_ = [("threshold", C.c_double), ("pair_cutoff", C.c_double), ("device", C.c_int), ("rank", C.c_int),
                ("world_size", C.c_int), ("verbose", C.c_int), ("j_two_limb", C.c_int), ("reserved", C.c_int * 7)]


class CfStats(C.Structure):
    _fields_ = [("nshell", C.c_int), ("nbf", C.c_int), ("ncart", C.c_int),
                ("shell_pairs_total", C.c_int64), ("shell_pairs_kept", C.c_int64),
                ("canonical_quartets", C.c_int64), ("canonical_quartets_local", C.c_int64),
                ("unique_integrals", C.c_int64), ("primitive_quartets", C.c_int64),
                ("flops_alg_jk", C.c_double * 4), ("n_launches_last", C.c_int),
                ("ms_device_last", C.c_double), ("ms_eri_last", C.c_double), ("fixedpoint_scale_log2", C.c_double * 2),
                ("threshold_effective_last", C.c_double), ("quartets_evaluated_last", C.c_int64),
                ("flops_alg_grad", C.c_double), ("ms_grad_last", C.c_double),
                ("primitive_quartets_executed_last", C.c_int64), ("flops_executed_last", C.c_double),
                ("j_two_limb_last", C.c_int), ("j_rounding_estimate_last", C.c_double),
                ("ref_repulsion_length", C.c_int64), ("ref_shell_quartet_length", C.c_int64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


_lib = None


def load_library(path: str = LIB_PATH):
    """Load libchinium_fock.so; raise (never fall back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FockEngineError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C chinium_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.c_int
    L.cf_create.restype = vp
    L.cf_create.argtypes = [C.POINTER(_CfBasis), C.POINTER(_CfOptions)]
    L.cf_create_multi.restype = vp
    L.cf_create_multi.argtypes = [C.POINTER(_CfBasis), C.POINTER(_CfOptions), ip, C.POINTER(C.c_int)]
    L.cf_num_devices.argtypes = [vp]
    L.cf_destroy.argtypes = [vp]
    L.cf_last_error.restype = C.c_char_p
    L.cf_last_error.argtypes = [vp]
    L.cf_get_stats.argtypes = [vp, C.POINTER(CfStats)]
    L.cf_nbf.argtypes = [vp]
    L.cf_set_density_threshold.argtypes = [vp, C.c_double]
    L.cf_get_repulsion_diag.argtypes = [vp, dp]
    L.cf_build_jk.argtypes = [vp, ip, dp, dp, dp, C.c_double, dp, dp, dp, dp]
    L.cf_build_jk_device.argtypes = [vp, ip, vp, vp, vp, C.c_double, vp, vp, vp, vp, vp]
    L.cf_acc_len.restype = C.c_size_t
    L.cf_acc_len.argtypes = [vp, ip]
    L.cf_acc_reduce_len.restype = C.c_size_t
    L.cf_acc_reduce_len.argtypes = [vp, ip]
    L.cf_accumulate_device.argtypes = [vp, ip, vp, vp, vp, C.c_double, vp, vp]
    L.cf_finalize_device.argtypes = [vp, ip, vp, C.c_double, ip, ip, ip, vp, vp, vp, vp, vp]
    L.cf_build_g_multi.argtypes = [vp, ip, ip, dp, C.c_double, dp]
    L.cf_contract_grads.argtypes = [vp, ip, dp, dp, C.c_double, ip, dp]
    L.cf_contract_grads_matrices.argtypes = [vp, ip, dp, C.c_double, ip, dp]
    if not (os.environ.get("CHINIUM_FOCK_LIB") and not hasattr(L, "cf_contract_hess")):   # A/B builds of older kernels may lack it
        L.cf_contract_hess.argtypes = [vp, ip, dp, C.c_double, ip, dp]
    L.cf_device_info.argtypes = [ip, C.c_char_p, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]
    L.cf_measure_fp64_peak.argtypes = [ip, dp]
    L.cf_sync_stats.argtypes = [vp]
    L.cf_one_electron.argtypes = [vp, ip, dp, dp, dp, dp, dp]
    L.cf_one_electron_device.argtypes = [vp, ip, dp, dp, vp, vp, vp, vp]
    L.cf_profile_tasks.argtypes = [vp, ip, vp, vp, vp, C.c_double, dp, ip, C.POINTER(ip)]
    _lib = L
    return L


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _fmat(a, nbf):
    if a is None:
        return None
    a = np.asfortranarray(a, dtype=np.float64)
    if a.size == 0:
        return None  # the reference's 0x0 "absent" matrix (Int4C2E.cpp:608-610)
    if a.shape != (nbf, nbf):
        raise FockEngineError(f"density has shape {a.shape}, expected ({nbf},{nbf})")
    return a


class Int4C2E:
    """Drop-in mirror of the reference class; `basis` is a chinium_b200.inputs.FlatBasis (the Mwfn stand-in)."""

    def __init__(self, basis, exx: float = 1.0, threshold: float = -1.0, device: int = -1, rank: int = 0,
                 world_size: int = 1, pair_cutoff: float = 0.0, j_two_limb: int = 0, ndevices: int = 0):
        """ndevices: 0 = single-device handle (cf_create); n > 0 = n GPUs of the box driven by this ONE process, -1 = all of
        them (cf_create_multi: worker thread + stream per device, NCCL int64 all-reduce inside the library)."""
        self.MWFN = basis
        self.EXX = float(exx)            # read at contract time, like the reference (SelfConsistentField.cpp:48)
        self.Threshold = float(threshold)
        self._opts = dict(device=device, rank=rank, world_size=world_size, pair_cutoff=pair_cutoff, j_two_limb=j_two_limb,
                          ndevices=ndevices)
        self._h = None
        self._lib = load_library()
        self.RepulsionDiags = None
        self.RepulsionLength = None
        self.ShellQuartetLength = None
        self._stage = 0

    # ---- handle management -----------------------------------------------------------------------
    def _ensure(self, output=0):
        if self._h is not None:
            return
        fb = self.MWFN
        keep = dict(type=np.ascontiguousarray(fb.type, np.int32), nprim=np.ascontiguousarray(fb.nprim, np.int32),
                    off=np.ascontiguousarray(fb.prim_offset, np.int32), exps=np.ascontiguousarray(fb.exps, np.float64),
                    coefs=np.ascontiguousarray(fb.coefs_normalized, np.float64),
                    xyz=np.ascontiguousarray(fb.center_xyz, np.float64).reshape(-1),
                    s2a=np.ascontiguousarray(fb.shell2atom, np.int32))
        ipt = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        b = _CfBasis(fb.nshell, ipt(keep["type"]), ipt(keep["nprim"]), ipt(keep["off"]), _dptr(keep["exps"]),
                     _dptr(keep["coefs"]), _dptr(keep["xyz"]), ipt(keep["s2a"]))
        o = _CfOptions()
        o.threshold = self.Threshold
        o.pair_cutoff = self._opts["pair_cutoff"]
        o.device = self._opts["device"]
        o.rank = self._opts["rank"]
        o.world_size = self._opts["world_size"]
        o.verbose = int(output)
        o.j_two_limb = int(self._opts["j_two_limb"])
        nd = int(self._opts["ndevices"])
        h = self._lib.cf_create(C.byref(b), C.byref(o)) if nd == 0 else self._lib.cf_create_multi(C.byref(b), C.byref(o), max(nd, 0), None)
        if not h:
            raise FockEngineError(self._lib.cf_last_error(None).decode())
        self._h = h
        self.nbf = self._lib.cf_nbf(h)

    def _check(self, rc):
        if rc != 0:
            raise FockEngineError("chinium_fock error %d: %s" % (rc, self._lib.cf_last_error(self._h).decode()))

    def close(self):
        if self._h is not None:
            self._lib.cf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setDensityThreshold(self, dthr: float):
        """Density-weighted screening for incremental builds (extension; see cf_set_density_threshold). 0 = off."""
        self._ensure()
        self._check(self._lib.cf_set_density_threshold(self._h, float(dthr)))

    @property
    def stats(self) -> dict:
        self._ensure()
        s = CfStats()
        self._check(self._lib.cf_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # ---- the reference's five setup stages (SelfConsistentField.cpp:49-53) -----------------------
    def getRepulsionDiag(self, output=0):
        self._ensure(output)
        d = np.zeros((self.nbf, self.nbf), order="F")
        self._check(self._lib.cf_get_repulsion_diag(self._h, _dptr(d)))
        self.RepulsionDiags = (d, None)   # only Diag1212 is ever consumed (Int4C2E.cpp:515,544)
        self._stage = max(self._stage, 1)

    def getRepulsionLength(self, output=0):
        assert self.RepulsionDiags is not None, "Diagonal elements of repulsion integrals are missing!"  # Int4C2E.cpp:514
        st = self.stats
        # the reference's own counts (its loop nest and uniqueness predicate, Int4C2E.cpp:79-128); the engine's work metric
        # (canonical shell quartets it evaluates) is stats["canonical_quartets"]
        self.RepulsionLength = st["ref_repulsion_length"]
        self.ShellQuartetLength = st["ref_shell_quartet_length"]
        if output > 0:
            nb, nsh = st["nbf"], st["nshell"]
            print("Before screening: %d integrals and %d shell quartets" % (
                nb * (nb + 1) * (nb * (nb + 1) // 2 + 1) // 4, nsh * (nsh + 1) * (nsh * (nsh + 1) // 2 + 1) // 4))
            print("After screening: %d integrals and %d shell quartets" % (self.RepulsionLength, self.ShellQuartetLength))
            print("Memory needed for 4c-2e repulsion integrals and their indices: 0 GB (direct build; the reference "
                  "would need %f GB)" % (self.RepulsionLength * 16.0 / 1024 ** 3))
        self._stage = max(self._stage, 2)

    def getRepulsionIndices(self, output=0):
        assert self._stage >= 2, "Shell quartet counts are missing!"
        self._stage = max(self._stage, 3)   # quartet lists are implicit (pair x pair tiles), nothing to store

    def getThreadPointers(self, nthreads=1, output=0):
        assert self._stage >= 3, "Shell indices are missing!"                               # Int4C2E.cpp:555
        self._stage = max(self._stage, 4)   # the static partition is (rank, world_size) of the handle

    def CalculateIntegrals(self, order=0, output=0):
        if order != 0:
            raise FockEngineError("derivative ERIs (order %d) are outside this engine's scope" % order)
        self._ensure(output)
        self._stage = max(self._stage, 5)

    # ---- the hot call -------------------------------------------------------------------------------
    def ContractInts(self, *args, nthreads=1, output=0):
        """ContractInts(Dd, Da, Db, nthreads=1, output=0) or ContractInts([D...], nthreads=1, output=0).
        `nthreads` is accepted and ignored (the GPU partition is the handle's), `output` > 0 prints the reference's line."""
        if len(args) >= 1 and isinstance(args[0], (list, tuple)):
            return self._contract_multi(list(args[0]))
        Dd, Da, Db = (list(args) + [None, None, None])[:3]
        return self._contract(Dd, Da, Db)

    def _contract(self, Dd, Da, Db, out=None):
        """`out`: optional (J, Kd, Ka, Kb) of caller-owned F-ordered nbf x nbf arrays to write into (e.g. views of pinned
        host memory); entries for absent densities may be None."""
        self._ensure()
        n = self.nbf
        Dd, Da, Db = _fmat(Dd, n), _fmat(Da, n), _fmat(Db, n)
        if out is not None:
            J = out[0]
            Ks = [out[1 + k] if D is not None else None for k, D in enumerate((Dd, Da, Db))]
            for a in [J] + [k for k in Ks if k is not None]:
                if a.shape != (n, n) or a.dtype != np.float64 or not a.flags["F_CONTIGUOUS"]:
                    raise FockEngineError("out matrices must be F-ordered float64 nbf x nbf")
            self._check(self._lib.cf_build_jk(self._h, n, _dptr(Dd), _dptr(Da), _dptr(Db), self.EXX, _dptr(J),
                                              _dptr(Ks[0]), _dptr(Ks[1]), _dptr(Ks[2])))
            return J, Ks[0], Ks[1], Ks[2]
        J = np.zeros((n, n), order="F")
        Ks = [np.zeros((n, n), order="F") if D is not None else None for D in (Dd, Da, Db)]
        self._check(self._lib.cf_build_jk(self._h, n, _dptr(Dd), _dptr(Da), _dptr(Db), self.EXX, _dptr(J),
                                          _dptr(Ks[0]), _dptr(Ks[1]), _dptr(Ks[2])))
        # the reference returns all-zero nbf x nbf matrices for absent K's (Int4C2E.cpp:616-619)
        Ks = [k if k is not None else np.zeros((n, n), order="F") for k in Ks]
        return J, Ks[0], Ks[1], Ks[2]

    def _contract_multi(self, Ds):
        self._ensure()
        n = self.nbf
        if len(Ds) == 0:
            return []
        Ds = [_fmat(D, n) for D in Ds]       # shape / dtype check: the C side reads nbf x nbf doubles per matrix
        if any(D is None for D in Ds):
            raise FockEngineError("ContractInts([D...]): every matrix must be nbf x nbf (None / 0x0 is not allowed here)")
        stack = np.ascontiguousarray(np.stack([D.T for D in Ds]))  # each block col-major
        out = np.zeros_like(stack)
        self._check(self._lib.cf_build_g_multi(self._h, n, len(Ds), _dptr(stack), self.EXX, _dptr(out)))
        return [np.asfortranarray(out[k].T) for k in range(len(Ds))]

    # ---- nuclear gradient (Int4C2E.cpp:747-763) ------------------------------------------------------
    def ContractGrads(self, *args):
        """The reference's three overloads (Int4C2E.h:45-47, Int4C2E.cpp:747-790):
          ContractGrads(D, output=0)            -> list of 3*natoms matrices G^(atom,xyz)[D] = d/dR (J[2D] - EXX K[D]); cached
                                                   in GradCache like the reference (:769-772)
          ContractGrads(D1, D2, output=0)       -> [3*natoms]: sum_ij D1_ij G^(.)[D2]_ij (fused on the device: the matrices
                                                   are never formed)
          ContractGrads([D1...], D2, output=0)  -> list of [3*natoms] vectors, one per D1 (reference: D1 o ContractGrads(D2))
        With world_size > 1 every form returns the partition's share: sum over the ranks."""
        if len(args) >= 1 and isinstance(args[0], (list, tuple)):
            D1s, D2 = args[0], args[1]
            Gs = self.ContractGrads(D2, 0)
            n = self.nbf
            return [np.array([float(np.sum(_fmat(D1, n) * G)) for G in Gs]) for D1 in D1s]
        if len(args) == 1 or (len(args) == 2 and np.ndim(args[1]) == 0):
            return self._contract_grads_matrices(args[0])
        return self._contract_grads(args[0], args[1])

    def _contract_grads_matrices(self, D):
        self._ensure()
        n = self.nbf
        D = _fmat(D, n)
        if D is None:
            raise FockEngineError("ContractGrads(D) needs an nbf x nbf matrix")
        if not hasattr(self, "GradCache"):
            self.GradCache = []
        for key, exx, value in self.GradCache:                 # the reference's isApprox lookup (Int4C2E.cpp:770)
            if exx == self.EXX and np.allclose(key, D, rtol=1e-12, atol=0.0):
                return value
        natom = int(np.max(self.MWFN.shell2atom)) + 1
        G = np.zeros((3 * natom, n * n))
        self._check(self._lib.cf_contract_grads_matrices(self._h, n, _dptr(D), self.EXX, natom, _dptr(G)))
        value = [np.asfortranarray(g.reshape(n, n).T) for g in G]
        self.GradCache.append((D.copy(), self.EXX, value))
        return value

    def _contract_grads(self, D1, D2):
        """ContractGrads(D1, D2, output) -> [3*natoms]: sum_ij D1_ij d/dR (J[2 D2] - EXX K[D2])_ij, index 3*atom + xyz."""
        self._ensure()
        n = self.nbf
        D1, D2 = _fmat(D1, n), _fmat(D2, n)
        if D1 is None or D2 is None:
            raise FockEngineError("ContractGrads needs two nbf x nbf matrices")
        natom = int(np.max(self.MWFN.shell2atom)) + 1
        g = np.zeros(3 * natom)
        self._check(self._lib.cf_contract_grads(self._h, n, _dptr(D1), _dptr(D2), self.EXX, natom, _dptr(g)))
        return g

    # ---- nuclear Hessian (Int4C2E.cpp:792-811) -------------------------------------------------------
    def ContractHesss(self, D1, D2, output=0):
        """The reference's ContractHesss(D1, D2, output) (Int4C2E.h:48, Int4C2E.cpp:792-811 over getRepulsion2 :410-492):
        3*natoms x 3*natoms matrix of sum over unique ERIs of d^2(ab|cd)/dX dY [2 D_ab D_cd - EXX/2 (D_ac D_bd + D_ad D_bc)].
        Like the reference (:793: `D = D1; D = D2;`) only D2 is used.  With world_size > 1: the partition's share."""
        self._ensure()
        n = self.nbf
        if output > 0:
            print("Contracting 4c-2e repulsion integral nuclear hessian with 1 matrix ... ", end="")
        t0 = time.perf_counter()
        D = _fmat(D2, n)
        if D is None:
            raise FockEngineError("ContractHesss needs an nbf x nbf matrix")
        natom = int(np.max(self.MWFN.shell2atom)) + 1
        Hm = np.zeros((3 * natom, 3 * natom), order="F")
        self._check(self._lib.cf_contract_hess(self._h, n, _dptr(D), self.EXX, natom, _dptr(Hm)))
        if output > 0:
            print("Done in %f s" % (time.perf_counter() - t0))
        return Hm

    # ---- device-resident / multi-GPU building blocks (plain pointers: e.g. torch tensors' data_ptr()) ---
    def acc_len(self, nk):
        """int64 words to allocate for an accumulator: [J | K.. | J low limb | scale tail]"""
        self._ensure()
        return int(self._lib.cf_acc_len(self._h, nk))

    def acc_reduce_len(self, nk):
        """leading words that a multi-GPU caller all-reduces (everything but the scale tail)"""
        self._ensure()
        return int(self._lib.cf_acc_reduce_len(self._h, nk))

    def accumulate_device(self, Dd_ptr, Da_ptr, Db_ptr, acc_ptr, stream=None):
        self._ensure()
        self._check(self._lib.cf_accumulate_device(self._h, self.nbf, Dd_ptr, Da_ptr, Db_ptr, self.EXX, acc_ptr, stream))

    def finalize_device(self, acc_ptr, has, J_ptr, Kd_ptr, Ka_ptr, Kb_ptr, stream=None):
        self._ensure()
        self._check(self._lib.cf_finalize_device(self._h, self.nbf, acc_ptr, self.EXX, int(has[0]), int(has[1]), int(has[2]),
                                                 J_ptr, Kd_ptr, Ka_ptr, Kb_ptr, stream))

    def build_jk_device(self, Dd_ptr, Da_ptr, Db_ptr, J_ptr, Kd_ptr, Ka_ptr, Kb_ptr, stream=None):
        self._ensure()
        self._check(self._lib.cf_build_jk_device(self._h, self.nbf, Dd_ptr, Da_ptr, Db_ptr, self.EXX, J_ptr, Kd_ptr, Ka_ptr,
                                                 Kb_ptr, stream))

    # ---- one-electron integrals (Int2C1E.cpp:18-67, :313-333) -----------------------------------
    def one_electron(self, Z, xyz_bohr):
        """-> (S, T, V) host matrices: overlap, kinetic energy, nuclear attraction of the point charges Z at xyz_bohr."""
        self._ensure()
        n = self.nbf
        Zd = np.ascontiguousarray(Z, np.float64)
        R = np.ascontiguousarray(xyz_bohr, np.float64).reshape(-1)
        S, T, V = (np.zeros((n, n), order="F") for _ in range(3))
        self._check(self._lib.cf_one_electron(self._h, len(Zd), _dptr(Zd), _dptr(R), _dptr(S), _dptr(T), _dptr(V)))
        return S, T, V

    def one_electron_device(self, Z, xyz_bohr, S_ptr, T_ptr, V_ptr, stream=None):
        """Same into DEVICE matrices (plain pointers); enqueued on `stream`, not synchronised."""
        self._ensure()
        Zd = np.ascontiguousarray(Z, np.float64)
        R = np.ascontiguousarray(xyz_bohr, np.float64).reshape(-1)
        self._check(self._lib.cf_one_electron_device(self._h, len(Zd), _dptr(Zd), _dptr(R), S_ptr, T_ptr, V_ptr, stream))

    def profile_tasks(self, Dd_ptr, Da_ptr, Db_ptr):
        """-> list of dicts (bra, ket class names, quartets, ms, flops_alg, group): every class-pair kernel timed alone."""
        self._ensure()
        rows = np.zeros((64, 8))
        n = C.c_int(0)
        self._check(self._lib.cf_profile_tasks(self._h, self.nbf, Dd_ptr, Da_ptr, Db_ptr, self.EXX, _dptr(rows), 64, C.byref(n)))
        names = ["ss", "ps", "pp", "ds", "dp", "dd", "fs", "fp", "fd", "ff"]
        return [dict(bra=names[int(r[0])], ket=names[int(r[1])], quartets=int(r[2]), ms=float(r[3]), flops_alg=float(r[4]),
                     group=int(r[5]), prim_executed=int(r[6]), flops_executed=float(r[7])) for r in rows[:n.value]]

    def sync_stats(self):
        self._check(self._lib.cf_sync_stats(self._h))   # also the deferred fixed-point range check of that build
        return self.stats


class Int2C1E:
    """Mirror of the reference's `class Int2C1E` (src/Integral/Int2C1E.h:10-57) for the zeroth-order matrices the SCF needs:
    CalculateIntegrals(0, output) fills Overlap, Kinetic, Nuclear (Int2C1E.cpp:313-333) on the GPU.  Multipole matrices,
    ECP terms and the derivative orders are outside this engine's scope.  `engine`: an Int4C2E on the same basis whose
    device handle is shared (otherwise one is created)."""

    def __init__(self, basis, Z, xyz_bohr, engine=None, device: int = -1):
        self.MWFN = basis
        self._Z = np.ascontiguousarray(Z, np.float64)
        self._xyz = np.ascontiguousarray(xyz_bohr, np.float64)
        self._eng = engine if engine is not None else Int4C2E(basis, 1.0, -1.0, device=device)
        self.Overlap = self.Kinetic = self.Nuclear = None

    def CalculateIntegrals(self, order=0, output=0):
        if order != 0:
            raise FockEngineError("one-electron derivative integrals (order %d) are outside this engine's scope" % order)
        import time
        t = time.perf_counter()
        if output > 0:
            print("Calculating 2c-1e integrals ... ", end="")
        self.Overlap, self.Kinetic, self.Nuclear = self._eng.one_electron(self._Z, self._xyz)
        if output > 0:
            print("Done in %f s" % (time.perf_counter() - t))


def measure_fp64_peak(device=-1) -> float:
    L = load_library()
    v = C.c_double(0)
    rc = L.cf_measure_fp64_peak(device, C.byref(v))
    if rc != 0:
        raise FockEngineError("cf_measure_fp64_peak failed: %s" % L.cf_last_error(None).decode())
    return v.value


def device_info(device=-1):
    L = load_library()
    name = C.create_string_buffer(128)
    sm, ma, mi = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.cf_device_info(device, name, 128, C.byref(sm), C.byref(ma), C.byref(mi))
    if rc != 0:
        raise FockEngineError("no CUDA device: %s" % L.cf_last_error(None).decode())
    return dict(name=name.value.decode(), sm_count=sm.value, cc=(ma.value, mi.value))
