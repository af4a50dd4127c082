// adaptor_test.cpp -- exercises chinium_b200/cpp/Int4C2E_b200.hpp the way the reference's SCF driver uses
// Int4C2E (src/HartreeFockKohnSham/SelfConsistentField.cpp:47-53, Restricted/SP.cpp:47, Restricted/Grad.cpp:66), with a column-major
// matrix shim standing in for Eigen::MatrixXd.  Input: a flat text dump of the basis and a density written by
// tests/test_gpu.py; output: J and K as text.  Exit code 3 = "no GPU" (the loud failure the CPU test checks).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../../chinium_b200/cpp/Int4C2E_b200.hpp"

struct Mat {   // the subset of Eigen::MatrixXd the adaptor touches
    long r = 0, c = 0;
    std::vector<double> v;
    Mat() {}
    Mat(long rows, long cols) : r(rows), c(cols), v((size_t)rows * cols) {}
    double* data() { return v.data(); }
    const double* data() const { return v.data(); }
    long rows() const { return r; }
    long cols() const { return c; }
    long size() const { return r * c; }
};
using Int4C2E = chinium_b200::Int4C2E_T<Mat>;

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: adaptor_test <basis+density.txt> <out.txt>\n"); return 2; }
    std::ifstream in(argv[1]);
    int nshell, nbf;
    in >> nshell >> nbf;
    chinium_b200::FlatBasis fb;
    for (int s = 0; s < nshell; s++) {
        int type, np, atom; double xyz[3];
        in >> type >> np >> atom >> xyz[0] >> xyz[1] >> xyz[2];
        std::vector<double> e(np), c(np);
        for (int k = 0; k < np; k++) in >> e[k] >> c[k];
        fb.add_shell(type, e, c, xyz, atom);
    }
    Mat D(nbf, nbf);
    for (long i = 0; i < D.size(); i++) in >> D.v[i];
    try {
        Int4C2E int4c2e;
        int4c2e = Int4C2E(fb, 1, -1);          // copy-assigned, like the reference
        int4c2e.EXX = 0.5;                     // overwritten after construction
        int4c2e.getRepulsionDiag(1);
        int4c2e.getRepulsionLength(1);
        int4c2e.getRepulsionIndices(1);
        int4c2e.getThreadPointers(4, 1);
        int4c2e.CalculateIntegrals(0, 1);
        Int4C2E copy = int4c2e;                // shares the device handle
        auto [J, Kd, Ka, Kb] = copy.ContractInts(D, Mat(0, 0), Mat(0, 0), 4, 1);
        std::vector<Mat> Ds{D, D};
        auto Gs = copy.ContractInts(Ds, 4, 1);
        std::ofstream out(argv[2]);
        out.precision(17);
        out << int4c2e.RepulsionLength << " " << int4c2e.ShellQuartetLength << "\n";
        for (double x : J.v) out << x << "\n";
        for (double x : Kd.v) out << x << "\n";
        for (double x : Gs[1].v) out << x << "\n";
        double ka = 0; for (double x : Ka.v) ka += x * x;
        out << ka << "\n";
        const std::vector<double> grads = copy.ContractGrads(D, D, 1);     // Restricted/Grad.cpp:66
        out << grads.size() << "\n";
        for (double x : grads) out << x << "\n";
        const std::vector<std::vector<double>> hess = copy.ContractHesss(D, D, 1);   // Restricted/Hess.cpp:67
        out << "HESS " << hess.size() << "\n";
        for (const auto& row : hess) for (double x : row) out << x << "\n";
        if (argc > 3) {   // the same calls through ONE process driving several GPUs (cf_create_multi): must be bit-identical
            const int nd = std::atoi(argv[3]);
            Int4C2E multi = Int4C2E::MultiDevice(fb, 1, -1, nd);
            multi.EXX = 0.5;
            multi.getRepulsionDiag(0); multi.getRepulsionLength(0); multi.getRepulsionIndices(0); multi.getThreadPointers(4, 0);
            multi.CalculateIntegrals(0, 0);
            auto [J2, Kd2, Ka2, Kb2] = multi.ContractInts(D, Mat(0, 0), Mat(0, 0), 4, 1);
            auto Gs2 = multi.ContractInts(Ds, 4, 1);
            const std::vector<double> grads2 = multi.ContractGrads(D, D, 0);
            const std::vector<std::vector<double>> hess2 = multi.ContractHesss(D, D, 0);
            bool same = J2.v == J.v && Kd2.v == Kd.v && Gs2[1].v == Gs[1].v && multi.RepulsionLength == int4c2e.RepulsionLength;
            double dg = 0; for (size_t i = 0; i < grads.size(); i++) dg = std::max(dg, std::fabs(grads[i] - grads2[i]));
            double hmax = 1.0;
            for (size_t i = 0; i < hess.size(); i++) for (size_t j = 0; j < hess.size(); j++) hmax = std::max(hmax, std::fabs(hess[i][j]));
            for (size_t i = 0; i < hess.size(); i++) for (size_t j = 0; j < hess.size(); j++) dg = std::max(dg, std::fabs(hess[i][j] - hess2[i][j]) / hmax);
            std::printf("multi-device handle: %d GPUs, J/K/G bit-identical to one GPU: %s, max|dgrad| %.3e\n", multi.NumDevices(), same ? "yes" : "NO", dg);
            out << "MULTI " << multi.NumDevices() << " " << (same ? 1 : 0) << " " << dg << "\n";
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "adaptor_test: %s\n", e.what());
        return std::strstr(e.what(), "no CUDA device") || std::strstr(e.what(), "no CPU fallback") ? 3 : 1;
    }
    return 0;
}
