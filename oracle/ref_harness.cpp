// C entry points around the UNMODIFIED reference sources.  TEST INFRASTRUCTURE ONLY.
//
// Built by oracle/Makefile from /root/reference/src/{definitions,lattice,ising,mcrg,rgnn}.cpp (compiled where
// they lie; nothing is copied into this repo) against oracle/mpi_stub/mpi.h, into oracle/_ref/libmcrg_ref.so.
// Every function below only *calls* reference code; the one loop that is restated here
// (ref_mcrg_loop, the body of mcrg.cpp:72-98, which is buried inside calc_critical_exponent) is
// checked against the real calc_critical_exponent by tests/test_oracle_vs_ref.py (same seed => same lambdas).
//
// Spin arrays cross this boundary in the reference's own layout: int32, column-major, element (i,j) at
// j*N+i (definitions.hpp:16).

// std / Eigen / stub headers first, so that the `private` override below touches reference headers only.
#include <random>
#include <cmath>
#include <iostream>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <vector>
#include <memory>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdio.h>
#include <unistd.h>
#include "Eigen/Dense"
#include <mpi.h>

#define private public
#include "definitions.hpp"
#include "lattice.hpp"
#include "ising.hpp"
#include "mcrg.hpp"
#include "rgnn.hpp"
#undef private

namespace {

imat to_imat(int N, const int *spins) {
    imat m(N, N);
    std::memcpy(m.data(), spins, sizeof(int) * (size_t)N * (size_t)N);
    return m;
}

void from_imat(const imat &m, int *out) {
    std::memcpy(out, m.data(), sizeof(int) * (size_t)m.rows() * (size_t)m.cols());
}

// Silence the constructor banner of MonteCarloRenormalizationGroup (mcrg.cpp:10-18) while it is built.
struct QuietStdout {
    int saved;
    QuietStdout() {
        fflush(stdout);
        saved = dup(1);
        FILE *nul = fopen("/dev/null", "w");
        dup2(fileno(nul), 1);
        fclose(nul);
    }
    ~QuietStdout() {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

}  // namespace

extern "C" {

// Reseed the reference's process-global mt19937_64 (definitions.hpp:28, definitions.cpp:4).
void ref_seed(unsigned long long s) { rng.seed(s); }

// Lattice::calc_interactions (lattice.cpp:102-120): out = {S_nn, S_nnn}, double-counted sums.
void ref_calc_interactions(int N, const int *spins, double *out2) {
    Lattice lat(1, to_imat(N, spins));
    vec2D S = lat.calc_interactions();
    out2[0] = S(0);
    out2[1] = S(1);
}

// Lattice::calc_nearest_neighbor_interaction (lattice.cpp:84-99).
double ref_calc_nn(int N, const int *spins) {
    Lattice lat(1, to_imat(N, spins));
    return lat.calc_nearest_neighbor_interaction();
}

// IsingModel::calc_energy (ising.cpp:158-173): K*S_nn/N^2 accumulated term by term.
double ref_calc_energy(int N, const int *spins, double K) {
    std::shared_ptr<Lattice> lat(new Lattice(1, to_imat(N, spins)));
    IsingModel ising(K);
    return ising.calc_energy(lat);
}

// IsingModel::calc_magnetization (ising.cpp:176-179): INTEGER division sum/N^2.
double ref_calc_magnetization(int N, const int *spins, double K) {
    std::shared_ptr<Lattice> lat(new Lattice(1, to_imat(N, spins)));
    IsingModel ising(K);
    return ising.calc_magnetization(lat);
}

// MonteCarloRenormalizationGroup::block_spin_transformation (mcrg.cpp:314-348, private).
// out is (N/b)x(N/b); ties are broken by the reference's global rng (seed it first).  Returns child a_.
int ref_block_spin(int N, int b, int a, const int *spins, int *out) {
    std::unique_ptr<MonteCarloRenormalizationGroup> rg;
    {
        QuietStdout q;
        rg.reset(new MonteCarloRenormalizationGroup(b));
    }
    std::shared_ptr<Lattice> lat(new Lattice(a, to_imat(N, spins)));
    std::shared_ptr<Lattice> child = rg->block_spin_transformation(lat);
    from_imat(child->spins_, out);
    return child->a_;
}

// Lattice::Lattice(int N) hot start (lattice.cpp:3-15, 33-41).
void ref_hot_lattice(int N, int *out) {
    Lattice lat(N);
    from_imat(lat.spins_, out);
}

// n Wolff updates via IsingModel::sample_new_configuration (ising.cpp:87-93) on a caller-supplied lattice.
void ref_wolff(int N, int *spins, double K, int n_updates) {
    std::shared_ptr<Lattice> lat(new Lattice(1, to_imat(N, spins)));
    IsingModel ising(K);
    for (int n = 0; n < n_updates; ++n) ising.sample_new_configuration(lat);
    from_imat(lat->spins_, spins);
}

// nearest_neighbors / next_nearest_neighbors tables (lattice.cpp:124-151): out is 4x2 column-major.
void ref_neighbors(int N, int i, int j, int *nn8, int *nnn8) {
    Lattice lat(1, imat::Ones(N, N));
    imat a = lat.nearest_neighbors(i, j);
    imat b = lat.next_nearest_neighbors(i, j);
    std::memcpy(nn8, a.data(), 8 * sizeof(int));
    std::memcpy(nnn8, b.data(), 8 * sizeof(int));
}

int ref_split_samples(int rank, int n_processes, int n_samples) {
    return split_samples(rank, n_processes, n_samples);
}

// flatten (definitions.cpp:9-19) of a 2x2 given column-major in4 -> out4.
void ref_flatten2(const double *in4, double *out4) {
    mat M(2, 2);
    std::memcpy(M.data(), in4, 4 * sizeof(double));
    vec v = flatten(M);
    for (int k = 0; k < 4; ++k) out4[k] = v(k);
}

void ref_rounded_str(double num, int precision, char *buf, int buflen) {
    std::string s = get_rounded_str(num, precision);
    std::snprintf(buf, buflen, "%s", s.c_str());
}

int ref_write_iter(int i) { return write_iter(i) ? 1 : 0; }

// The real calc_critical_exponent (mcrg.cpp:22-144).  It writes critical_exponent_N_<N>_K_<K>.txt into the
// CWD; the caller passes a scratch directory.  Parses "n, lambda, nu" rows back; returns number of levels.
int ref_critical_exponent(const char *workdir, int n_eq, int n_samples, int N, double K, double *lambdas,
                          double *nus, int max_levels) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(workdir) != 0) return -2;
    {
        QuietStdout q;
        MonteCarloRenormalizationGroup rg(2);
        rg.calc_critical_exponent(n_eq, n_samples, N, K);
    }
    std::string filename = "critical_exponent_N_" + std::to_string(N) + "_K_" + get_rounded_str(K, 7) + ".txt";
    FILE *f = fopen(filename.c_str(), "r");
    int n_found = 0;
    if (f) {
        char line[512];
        while (fgets(line, sizeof line, f)) {
            if (line[0] == '#') continue;
            int n;
            double lam, nu;
            if (sscanf(line, " %d , %lf , %lf", &n, &lam, &nu) == 3 && n_found < max_levels) {
                lambdas[n_found] = lam;
                nus[n_found] = nu;
                ++n_found;
            }
        }
        fclose(f);
    }
    if (chdir(cwd) != 0) return -3;
    return n_found;
}

// The real locate_critical_point (mcrg.cpp:146-310, two-lattice matching with the NN operator).  It writes
// critical_point_L_<L>_K_<K0>.txt into the CWD; the caller passes a scratch directory.  Parses the rows
// "iteration, level, starting K, approximate Kc" back into out[(iteration-1)*n_levels + level] = Kc; returns the rows found.
int ref_locate_critical_point(const char *workdir, int n_iterations, int n_eq, int n_samples, int L, double K0, double *out,
                              int max_rows, double *K_final) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(workdir) != 0) return -2;
    {
        QuietStdout q;
        MonteCarloRenormalizationGroup rg(2);
        const double K = rg.locate_critical_point(n_iterations, n_eq, n_samples, L, K0);
        if (K_final) *K_final = K;
    }
    std::string filename = "critical_point_L_" + std::to_string(L) + "_K_" + get_rounded_str(K0, 7) + ".txt";
    FILE *f = fopen(filename.c_str(), "r");
    int n_found = 0;
    if (f) {
        char line[512];
        while (fgets(line, sizeof line, f)) {
            if (line[0] == '#') continue;
            int it, lv;
            double Ks, Kc;
            if (sscanf(line, " %d , %d , %lf , %lf", &it, &lv, &Ks, &Kc) == 4 && n_found < max_rows) out[n_found++] = Kc;
        }
        fclose(f);
    }
    if (chdir(cwd) != 0) return -3;
    return n_found;
}

// The sample loop of calc_critical_exponent (mcrg.cpp:42-98) with the per-sample S matrix logged.
// Same construction order and rng consumption as the reference, so with the same seed it visits the same
// configurations as ref_critical_exponent.  S_log receives n_samples*(n_lv+1)*2 doubles
// ([sample][level][op]); pass NULL to only time it.  Returns elapsed seconds of the sampling loop.
double ref_mcrg_loop(int n_eq, int n_samples, int N, double K, int cold_start, double *S_log, int *n_levels_out) {
    std::unique_ptr<MonteCarloRenormalizationGroup> rg;
    {
        QuietStdout q;
        rg.reset(new MonteCarloRenormalizationGroup(2));
    }
    int n_transformations = floor(log(N) / log(2)) - 1;  // mcrg.cpp:43
    if (n_levels_out) *n_levels_out = n_transformations;
    std::unique_ptr<IsingModel> pIsing(new IsingModel(K));
    std::shared_ptr<Lattice> pLattice(new Lattice(N));
    if (cold_start) pLattice->spins_.setOnes();
    pIsing->equilibrate(pLattice, n_eq, false);
    std::shared_ptr<Lattice> pLatticeb;
    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < n_samples; ++s) {
        pIsing->sample_new_configuration(pLattice);
        vec2D S0 = pLattice->calc_interactions();
        if (S_log) {
            S_log[((size_t)s * (n_transformations + 1)) * 2 + 0] = S0(0);
            S_log[((size_t)s * (n_transformations + 1)) * 2 + 1] = S0(1);
        }
        pLatticeb = pLattice;
        for (int n = 1; n <= n_transformations; ++n) {
            pLatticeb = rg->block_spin_transformation(pLatticeb);
            vec2D Sn = pLatticeb->calc_interactions();
            if (S_log) {
                S_log[((size_t)s * (n_transformations + 1) + n) * 2 + 0] = Sn(0);
                S_log[((size_t)s * (n_transformations + 1) + n) * 2 + 1] = Sn(1);
            }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Equilibrium statistics from the reference's own sampler: after n_eq Wolff updates, n_samples measurements
// `stride` updates apart.  Logs per sample {S_nn, sum of spins} (exact integers in doubles) so that the
// caller forms <E>, <|M|>, U4 and their errors however it likes.
void ref_thermo_series(int N, double K, int n_eq, int n_samples, int stride, int cold_start, double *out2) {
    std::unique_ptr<IsingModel> pIsing(new IsingModel(K));
    std::shared_ptr<Lattice> pLattice(new Lattice(N));
    if (cold_start) pLattice->spins_.setOnes();
    pIsing->equilibrate(pLattice, n_eq, false);
    for (int s = 0; s < n_samples; ++s) {
        for (int k = 0; k < stride; ++k) pIsing->sample_new_configuration(pLattice);
        out2[2 * (size_t)s + 0] = pLattice->calc_nearest_neighbor_interaction();
        out2[2 * (size_t)s + 1] = (double)pLattice->spins_.sum();
    }
}

// RGNN forward pass scalar_output (rgnn.cpp:281-307) with caller-supplied b x b weights (column-major).
double ref_rgnn_scalar_output(int N, const int *spins, int b, const double *W) {
    std::unique_ptr<RenormalizationGroupNeuralNetwork> net(new RenormalizationGroupNeuralNetwork(b));
    mat Wm(b, b);
    std::memcpy(Wm.data(), W, sizeof(double) * b * b);
    net->set_weights(Wm);
    return net->scalar_output(to_imat(N, spins));
}

// RGNN finite-difference gradient (rgnn.cpp:310-339); grad out is b x b column-major.
void ref_rgnn_gradient(int N, const int *spins, int b, const double *W, double h, double *grad) {
    std::unique_ptr<RenormalizationGroupNeuralNetwork> net(new RenormalizationGroupNeuralNetwork(b));
    mat Wm(b, b);
    std::memcpy(Wm.data(), W, sizeof(double) * b * b);
    net->set_weights(Wm);
    mat g = net->calc_gradient_scalar_output(h, to_imat(N, spins));
    std::memcpy(grad, g.data(), sizeof(double) * b * b);
}

}  // extern "C"
