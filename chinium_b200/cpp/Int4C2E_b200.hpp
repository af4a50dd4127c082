// Int4C2E_b200.hpp -- C++ adaptor with the reference's `class Int4C2E` surface over the C ABI.
//
// The reference's boundary for the J/K path is the class `Int4C2E` (src/Integral/Int4C2E.h:10-50) as used by
// src/HartreeFockKohnSham/SelfConsistentField.cpp:47-53 (construction + five setup stages) and by
// Restricted/SP.cpp:47, Unrestricted/SP.cpp:58, Universal.cpp:38 (the per-iteration ContractInts).  This header
// keeps those method names, argument meaning and error behaviour and forwards everything numerical to
// libchinium_fock.so (include/chinium_fock.h).  Header-only; templated on the matrix type so that it compiles
//   * inside Chinium with  `using Int4C2E = chinium_b200::Int4C2E_T<EigenMatrix>;`  (Eigen::MatrixXd: col-major,
//     .data()/.rows()/.cols()/.size(), constructible as M(rows, cols)), and
//   * in this repository's own test (tests/cpp/adaptor_test.cpp) with a 30-line column-major shim,
// because Eigen and libmwfn are not available in the build image.
// The class is cheaply copyable (the reference copy-assigns it, SelfConsistentField.cpp:47): the device handle
// is held by a std::shared_ptr.  No CPU fallback: without a usable sm_100 GPU every call throws.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/chinium_fock.h"

namespace chinium_b200 {

// What __Make_Basis_Set__ (src/Integral/Macro.h:1-25) reads from an Mwfn, flattened.  `from_mwfn` below fills it
// from any object with the libmwfn layout (Centers[].Coordinates / Centers[].Shells[].{Type, Exponents,
// NormalizedCoefficients}); INTEGRATION.md shows the call.
struct FlatBasis {
    std::vector<int> type, nprim, prim_offset, shell2atom;
    std::vector<double> exps, coefs_normalized, center_xyz;
    int nshell() const { return (int)type.size(); }
    void add_shell(int type_, const std::vector<double>& e, const std::vector<double>& c, const double xyz[3], int atom) {
        type.push_back(type_);
        nprim.push_back((int)e.size());
        prim_offset.push_back((int)exps.size());
        exps.insert(exps.end(), e.begin(), e.end());
        coefs_normalized.insert(coefs_normalized.end(), c.begin(), c.end());
        center_xyz.insert(center_xyz.end(), xyz, xyz + 3);
        shell2atom.push_back(atom);
    }
    cf_basis view() const {
        cf_basis b;
        b.nshell = nshell();
        b.type = type.data(); b.nprim = nprim.data(); b.prim_offset = prim_offset.data();
        b.exps = exps.data(); b.coefs_normalized = coefs_normalized.data(); b.center_xyz = center_xyz.data();
        b.shell2atom = shell2atom.data();
        return b;
    }
};

template <class MwfnLike>
FlatBasis from_mwfn(const MwfnLike& mwfn) {   // shell order = centre order x shell order (Macro.h:9-24)
    FlatBasis fb;
    int atom = 0;
    for (const auto& center : mwfn.Centers) {
        const double xyz[3] = {center.Coordinates[0], center.Coordinates[1], center.Coordinates[2]};
        for (const auto& shell : center.Shells)
            fb.add_shell(shell.Type, shell.Exponents, shell.NormalizedCoefficients, xyz, atom);
        atom++;
    }
    return fb;
}

template <class Matrix>
class Int4C2E_T {
  public:
    double Threshold = -1;
    double EXX = 1;                                     // read at contract time (SelfConsistentField.cpp:48)
    std::tuple<Matrix, Matrix> RepulsionDiags;          // only Diag1212 is consumed (Int4C2E.cpp:515,544)
    long int ShellQuartetLength = 0;
    long int RepulsionLength = 0;

    Int4C2E_T() {}
    Int4C2E_T(const FlatBasis& basis, double exx, double threshold, int device = -1, int rank = 0, int world_size = 1)
        : Threshold(threshold), EXX(exx), basis_(std::make_shared<FlatBasis>(basis)), device_(device), rank_(rank), world_(world_size) {}
    // all-in-one-process multi-GPU: `ndevices` GPUs of the box (<= 0: all of them) behind the same ContractInts call;
    // partial J/K are summed with an integer NCCL all-reduce inside the library (cf_create_multi)
    static Int4C2E_T MultiDevice(const FlatBasis& basis, double exx, double threshold, int ndevices) {
        Int4C2E_T x(basis, exx, threshold);
        x.ndevices_ = ndevices <= 0 ? -1 : ndevices;
        return x;
    }
    int NumDevices() { ensure(0); return cf_num_devices(h_.get()); }

    // ---- the five setup stages, same order and same assertions as the reference (Int4C2E.cpp:500-587)
    void getRepulsionDiag(int output) {
        ensure(output);
        const int n = cf_nbf(h_.get());
        Matrix d(n, n);
        check(cf_get_repulsion_diag(h_.get(), d.data()));
        RepulsionDiags = std::make_tuple(d, Matrix(0, 0));
        stage_ = stage_ < 1 ? 1 : stage_;
    }
    void getRepulsionLength(int output) {
        if (stage_ < 1) throw std::runtime_error("Diagonal elements of repulsion integrals are missing!");   // Int4C2E.cpp:514
        cf_stats st;
        check(cf_get_stats(h_.get(), &st));
        RepulsionLength = (long int)st.ref_repulsion_length;         // the reference's own counts (its loop nest and uniqueness
        ShellQuartetLength = (long int)st.ref_shell_quartet_length;  // predicate, Int4C2E.cpp:79-128), not the engine's work metric
        if (output > 0) {
            const long nb = st.nbf, ns = st.nshell;
            std::printf("Before screening: %ld integrals and %ld shell quartets\n", nb * (nb + 1) * (nb * (nb + 1) / 2 + 1) / 4,
                        ns * (ns + 1) * (ns * (ns + 1) / 2 + 1) / 4);
            std::printf("After screening: %ld integrals and %ld shell quartets\n", RepulsionLength, ShellQuartetLength);
            std::printf("Memory needed for 4c-2e repulsion integrals and their indices: 0 GB (direct build on the GPU; "
                        "the stored list would need %f GB)\n", RepulsionLength * 16.0 / 1024 / 1024 / 1024);
        }
        stage_ = stage_ < 2 ? 2 : stage_;
    }
    void getRepulsionIndices(int /*output*/) {
        if (stage_ < 2) throw std::runtime_error("Shell quartet counts are missing!");                       // Int4C2E.cpp:538
        stage_ = stage_ < 3 ? 3 : stage_;    // quartet lists are implicit in the pair-class tiles on the device
    }
    void getThreadPointers(int /*nthreads*/, int /*output*/) {
        if (stage_ < 3) throw std::runtime_error("Shell indices are missing!");                              // Int4C2E.cpp:555
        stage_ = stage_ < 4 ? 4 : stage_;    // the static partition is (rank, world_size) of the handle
    }
    void CalculateIntegrals(int order, int output) {
        if (order != 0) throw std::runtime_error("derivative ERIs are outside the scope of the B200 J/K engine");
        ensure(output);
        stage_ = 5;
    }

    // ---- the hot call (Int4C2E.cpp:673-683): matrices by value, 0x0 = absent, four nbf x nbf matrices back
    std::tuple<Matrix, Matrix, Matrix, Matrix> ContractInts(Matrix Dd, Matrix Da, Matrix Db, int /*nthreads*/, int output) {
        ensure(0);
        const int n = cf_nbf(h_.get());
        auto t0 = std::chrono::steady_clock::now();
        if (output > 0) std::printf("Contracting 4c-2e repulsion integrals with 1 matrix ... ");
        auto in = [&](Matrix& m) -> const double* {
            if (m.size() == 0) return nullptr;
            if (m.rows() != n || m.cols() != n) throw std::runtime_error("ContractInts: density is not nbf x nbf");
            return m.data();
        };
        Matrix J(n, n), Kd(n, n), Ka(n, n), Kb(n, n);
        fill_zero(J); fill_zero(Kd); fill_zero(Ka); fill_zero(Kb);      // absent K's are all-zero nbf x nbf (Int4C2E.cpp:616-619)
        const double* pd = in(Dd); const double* pa = in(Da); const double* pb = in(Db);
        check(cf_build_jk(h_.get(), n, pd, pa, pb, EXX, J.data(), pd ? Kd.data() : nullptr, pa ? Ka.data() : nullptr,
                          pb ? Kb.data() : nullptr));
        if (output > 0) std::printf("Done in %f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return std::make_tuple(J, Kd, Ka, Kb);
    }
    // multi-density call (Int4C2E.cpp:685-745): G_k = J[2 D_k] - EXX K[D_k]
    std::vector<Matrix> ContractInts(std::vector<Matrix>& Ds, int /*nthreads*/, int output) {
        ensure(0);
        const int n = cf_nbf(h_.get());
        auto t0 = std::chrono::steady_clock::now();
        if (output > 0) std::printf("Contracting 4c-2e repulsion integrals with %d matrices ... ", (int)Ds.size());
        const size_t n2 = (size_t)n * n;
        std::vector<double> in(n2 * Ds.size()), out(n2 * Ds.size());
        for (size_t k = 0; k < Ds.size(); k++) {
            if (Ds[k].rows() != n || Ds[k].cols() != n) throw std::runtime_error("ContractInts: density is not nbf x nbf");
            std::copy(Ds[k].data(), Ds[k].data() + n2, in.begin() + k * n2);
        }
        if (!Ds.empty()) check(cf_build_g_multi(h_.get(), n, (int)Ds.size(), in.data(), EXX, out.data()));
        std::vector<Matrix> Gs;
        for (size_t k = 0; k < Ds.size(); k++) {
            Matrix G(n, n);
            std::copy(out.begin() + k * n2, out.begin() + (k + 1) * n2, G.data());
            Gs.push_back(G);
        }
        if (output > 0) std::printf("Done in %f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return Gs;
    }

    // nuclear gradient (Int4C2E.cpp:747-763): grad[3*atom + xyz] = sum D1 o d/dR (J[2 D2] - EXX K[D2]); the 3*natoms
    // intermediate matrices of getRepulsion1 (:312-408) are contracted on the fly and never formed
    std::vector<double> ContractGrads(Matrix D1, Matrix D2, int output) {
        ensure(0);
        const int n = cf_nbf(h_.get());
        if (D1.rows() != n || D1.cols() != n || D2.rows() != n || D2.cols() != n) throw std::runtime_error("ContractGrads: matrix is not nbf x nbf");
        auto t0 = std::chrono::steady_clock::now();
        if (output > 0) std::printf("Contracting 4c-2e repulsion integral nuclear gradient with 1 matrix from the left and 1 matrix from the right ... ");
        int natom = 0;
        for (int a : basis_->shell2atom) natom = a + 1 > natom ? a + 1 : natom;
        std::vector<double> g(3 * (size_t)natom, 0.0);
        check(cf_contract_grads(h_.get(), n, D1.data(), D2.data(), EXX, natom, g.data()));
        if (output > 0) std::printf("Done in %f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return g;
    }

    // matrix form (Int4C2E.cpp:766-790; consumer Restricted/Hess.cpp:72): 3*natoms matrices G^(atom,xyz)[D], kept in
    // GradCache and looked up by value like the reference does (:769-772)
    std::vector<std::tuple<Matrix, double, std::vector<Matrix>>> GradCache;
    std::vector<Matrix> ContractGrads(Matrix D, int output) {
        ensure(0);
        const int n = cf_nbf(h_.get());
        if (D.rows() != n || D.cols() != n) throw std::runtime_error("ContractGrads: matrix is not nbf x nbf");
        auto t0 = std::chrono::steady_clock::now();
        if (output > 0) std::printf("Contracting 4c-2e repulsion integral nuclear gradient with 1 matrix ... ");
        const size_t n2 = (size_t)n * n;
        for (auto& entry : GradCache) {
            Matrix& key = std::get<0>(entry);
            bool same = std::get<1>(entry) == EXX;
            for (size_t i = 0; same && i < n2; i++) same = std::fabs(key.data()[i] - D.data()[i]) <= 1e-12 * std::fabs(D.data()[i]);
            if (same) {
                if (output > 0) std::printf("Found in cache -> Done in %f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
                return std::get<2>(entry);
            }
        }
        int natom = 0;
        for (int a : basis_->shell2atom) natom = a + 1 > natom ? a + 1 : natom;
        std::vector<double> buf(3 * (size_t)natom * n2);
        check(cf_contract_grads_matrices(h_.get(), n, D.data(), EXX, natom, buf.data()));
        std::vector<Matrix> Gs;
        for (int j = 0; j < 3 * natom; j++) {
            Matrix G(n, n);
            std::copy(buf.begin() + (size_t)j * n2, buf.begin() + (size_t)(j + 1) * n2, G.data());
            Gs.push_back(G);
        }
        GradCache.push_back(std::make_tuple(D, EXX, Gs));
        if (output > 0) std::printf("Done in %f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return Gs;
    }
    // several left matrices against one right matrix (Int4C2E.cpp:752-764): D1s[i] o ContractGrads(D2)[j]
    std::vector<std::vector<double>> ContractGrads(std::vector<Matrix>& D1s, Matrix D2, int output) {
        const std::vector<Matrix> GD2 = ContractGrads(D2, output);
        const size_t n2 = (size_t)D2.rows() * D2.cols();
        std::vector<std::vector<double>> out(D1s.size(), std::vector<double>(GD2.size(), 0.0));
        for (size_t i = 0; i < D1s.size(); i++)
            for (size_t j = 0; j < GD2.size(); j++) {
                double s = 0.0;
                for (size_t k = 0; k < n2; k++) s += D1s[i].data()[k] * GD2[j].data()[k];
                out[i][j] = s;
            }
        return out;
    }

    // nuclear Hessian (Int4C2E.cpp:792-811 over getRepulsion2 :410-492; consumer Restricted/Hess.cpp:67): like the reference,
    // only D2 enters (`EigenMatrix D = D1; D = D2;`, :793); GDs[i][j], i, j = 3*atom + xyz
    std::vector<std::vector<double>> ContractHesss(Matrix D1, Matrix D2, int output) {
        (void)D1;
        ensure(0);
        const int n = cf_nbf(h_.get());
        if (D2.rows() != n || D2.cols() != n) throw std::runtime_error("ContractHesss: matrix is not nbf x nbf");
        auto t0 = std::chrono::steady_clock::now();
        if (output > 0) std::printf("Contracting 4c-2e repulsion integral nuclear hessian with 1 matrix ... ");
        int natom = 0;
        for (int a : basis_->shell2atom) natom = a + 1 > natom ? a + 1 : natom;
        const size_t nh = 3 * (size_t)natom;
        std::vector<double> buf(nh * nh, 0.0);
        check(cf_contract_hess(h_.get(), n, D2.data(), EXX, natom, buf.data()));
        std::vector<std::vector<double>> GDs(nh, std::vector<double>(nh));
        for (size_t i = 0; i < nh; i++) for (size_t j = 0; j < nh; j++) GDs[i][j] = buf[j * nh + i];
        if (output > 0) std::printf("Done in %f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return GDs;
    }

    // extension for direct SCF (no counterpart in the stored-ERI reference): density-weighted screening for incremental
    // builds G[D_n - D_(n-1)]; 0 switches it off (see cf_set_density_threshold)
    void setDensityThreshold(double dthr) { ensure(0); check(cf_set_density_threshold(h_.get(), dthr)); }

    cf_handle* handle() { ensure(0); return h_.get(); }

  private:
    std::shared_ptr<FlatBasis> basis_;
    std::shared_ptr<cf_handle> h_;
    int device_ = -1, rank_ = 0, world_ = 1, stage_ = 0;
    int ndevices_ = 0;      // 0: single-device handle (cf_create); -1: all GPUs; n > 0: n GPUs (cf_create_multi)

    static void fill_zero(Matrix& m) { double* p = m.data(); for (long i = 0; i < (long)m.size(); i++) p[i] = 0.0; }
    void ensure(int output) {
        if (h_) return;
        if (!basis_) throw std::runtime_error("Int4C2E: no basis (default-constructed object)");
        cf_options o{};
        o.threshold = Threshold; o.device = device_; o.rank = rank_; o.world_size = world_; o.verbose = output;
        cf_basis b = basis_->view();
        cf_handle* h = ndevices_ == 0 ? cf_create(&b, &o) : cf_create_multi(&b, &o, ndevices_ < 0 ? 0 : ndevices_, nullptr);
        if (!h) throw std::runtime_error(std::string("chinium_fock: ") + cf_last_error(nullptr));
        h_ = std::shared_ptr<cf_handle>(h, cf_destroy);
    }
    void check(int rc) {
        if (rc != CF_OK) throw std::runtime_error("chinium_fock error " + std::to_string(rc) + ": " + cf_last_error(h_.get()));
    }
};

}  // namespace chinium_b200
