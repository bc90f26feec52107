// Drop-in C++ interface of the MCRG hot path: the reference's class names, public members and signatures
// (lattice.hpp:5-37, ising.hpp:8-41, mcrg.hpp:6-40, definitions.hpp:11-50), implemented on top of the C ABI of
// libmcrg_b200.so (include/mcrg_b200.h).  A program written against the reference's headers — its own main.cpp —
// compiles unchanged against lattice.hpp / ising.hpp / mcrg.hpp / definitions.hpp in this directory, which simply
// include this file.
//
// What differs from the reference, by design (north_star):
//   * the Markov update is a checkerboard Metropolis sweep on the GPU, not a Wolff cluster flip (ising.cpp:87-155);
//     one "update" = MCRG_SWEEPS_PER_UPDATE sweeps (environment, default 1);
//   * the drivers (calc_critical_exponent, locate_critical_point) run MCRG_REPLICAS independent chains on the
//     device (default 1024) where the reference runs one chain per MPI rank; n_processes_ reports that number;
//   * errors are reported: a failing device call throws std::runtime_error (the reference checks nothing).
// Public data members keep their meaning.  Lattice::spins_ is host memory and is the source of truth for the
// fine-grained methods: each of them uploads it, works on the device and downloads the result.
#ifndef MCRG_B200_DROPIN_HPP
#define MCRG_B200_DROPIN_HPP

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <stdio.h>
#include <string>
#include <vector>

#include "Eigen/Dense"
#include "mpi.h"

using namespace Eigen;  // the reference's definitions.hpp:11 leaks this; programs written against it rely on it

// ---- definitions.hpp:15-20 ---------------------------------------------------------------------------------
using mat = Matrix<double, Dynamic, Dynamic, ColMajor>;
using imat = Matrix<int, Dynamic, Dynamic, ColMajor>;
using mat2D = Matrix<double, 2, 2>;
using vec = Matrix<double, Dynamic, 1>;
using ivec = Matrix<int, Dynamic, 1>;
using vec2D = Matrix<double, 2, 1>;

// definitions.hpp:22-23 / definitions.cpp:9-31: column-major flatten of a square matrix and its inverse
extern vec flatten(mat M);
extern mat unflatten(vec v);

// definitions.hpp:27-32: host-side generators (used for host-only helpers; device randomness is Philox)
extern std::random_device rd;
extern std::mt19937_64 rng;
extern std::uniform_int_distribution<int> binary;
extern std::uniform_real_distribution<double> unif;
extern std::normal_distribution<double> norm;

inline auto rand_spin() { return 2 * binary(rng) - 1; }
inline auto rand_weight() { return 0.3 * norm(rng); }
inline auto rand_unif() { return unif(rng); }

extern void display_spin_up();
extern void display_spin_down();
extern bool write_iter(int i);                                         // definitions.cpp:44-68 log-spaced schedule
extern std::string get_rounded_str(double num, int precision);         // definitions.cpp:70-77
extern int split_samples(int rank, int n_processes, int n_samples);    // definitions.cpp:79-87

namespace mcrg_b200 {
struct DeviceBatch;  // owns one mcrg_ctx
struct Settings {
    int replicas;           // MCRG_REPLICAS: chains per driver call
    int sweeps_per_update;  // MCRG_SWEEPS_PER_UPDATE: Metropolis sweeps standing in for one reference update
    int device;             // MCRG_DEVICE
    std::uint64_t seed;     // MCRG_SEED
    int quiet;              // MCRG_QUIET: suppress banners
    int cluster;            // MCRG_UPDATE: 1 = "cluster" (Swendsen-Wang updates, the reference's family, ising.cpp:87-155;
                            // sweeps_per_update then counts cluster updates), 0 = "metropolis", -1 = unset: cluster
                            // updates for N >= 32, Metropolis sweeps below
    int devices;            // MCRG_DEVICES: GPUs of this process that calc_critical_exponent spreads its chains over
                            // (devices 0..n-1; totals by one NCCL all-reduce, mcrg_allreduce_accumulators)
    int compat;             // MCRG_COMPAT (default 1): keep the reference's arithmetic in the thermodynamics log — the
                            // INTEGER division in calc_magnetization (ising.cpp:178) and the second division of the
                            // per-spin energy by n_spins (ising.cpp:72); 0 = |sum s| / N^2 as a real number, no second division
};
Settings &settings();
}  // namespace mcrg_b200

// ---- lattice.hpp:5-37 --------------------------------------------------------------------------------------
class Lattice {
public:
    Lattice(int N);
    Lattice(int a, imat spins);
    ~Lattice();

    int N_;
    int n_spins_;
    int a_;
    imat spins_;

    void write_spins(FILE *fptr);
    void display_spins();
    int choose_random_spin();
    double calc_nearest_neighbor_interaction();
    vec2D calc_interactions();
    imat nearest_neighbors(int i, int j);
    imat next_nearest_neighbors(int i, int j);

    // additions (not in the reference): the plaquette sum and the raw spin sum of the current spins_
    double calc_plaquette_interaction();
    long long sum_spins();
    std::shared_ptr<mcrg_b200::DeviceBatch> device_batch();  // 1-replica device mirror, created on first use

private:
    int rank_;
    int n_processes_;
    std::shared_ptr<mcrg_b200::DeviceBatch> dev_;
    std::uniform_int_distribution<int> pick_site_;
};

// ---- ising.hpp:8-41 ----------------------------------------------------------------------------------------
class IsingModel {
public:
    IsingModel(double K);
    ~IsingModel() {}

    void equilibrate(std::shared_ptr<Lattice>, int n_samples_eq, bool write);
    void sample_new_configuration(std::shared_ptr<Lattice> pLattice);
    double calc_magnetization(std::shared_ptr<Lattice> pLattice);
    double calc_energy(std::shared_ptr<Lattice> pLattice);

private:
    int rank_;
    int n_processes_;
    double K_;
    FILE *fptr_;
};

// ---- mcrg.hpp:6-40 -----------------------------------------------------------------------------------------
class MonteCarloRenormalizationGroup {
public:
    MonteCarloRenormalizationGroup(int b);
    ~MonteCarloRenormalizationGroup() {}

    int b_;

    void calc_critical_exponent(int n_samples_eq, int n_samples, int N, double K);
    double locate_critical_point(int n_iterations, int n_samples_eq, int n_samples, int L, double K0);

    // results of the last calc_critical_exponent call (additions; the reference only prints them)
    std::vector<double> lambdas_, nus_, lambda_errors_;
    // last approx_critical_point call: Kc per blocking level and its jackknife error over groups of chains
    std::vector<double> kcs_, kc_errors_;

private:
    int rank_;
    int n_processes_;
    int iter_;
    FILE *fptr_;

    double approx_critical_point(int n_samples_eq, int n_samples, int L, double K);
    std::shared_ptr<Lattice> block_spin_transformation(std::shared_ptr<Lattice> pLattice);
};

// ---- rgnn.hpp:7-66 -----------------------------------------------------------------------------------------
class RenormalizationGroupNeuralNetwork {
public:
    RenormalizationGroupNeuralNetwork(int b);
    ~RenormalizationGroupNeuralNetwork() {}

    int n_processes_;
    int rank_;
    int b_;
    int t_;
    double eta_;
    double beta1_;
    double beta2_;
    double epsilon_;
    double w_;
    mat m_;
    mat v_;
    mat W_;
    FILE *fptr_;

    double final_mse_;

    void initialize();
    void set_weights(const mat &W);
    void train_scalar_output(int L, int n_cycles, int n_samples, int n_samples_eq, double T, double h, double eta);
    void test_scalar_output(int L, int n_samples, int n_samples_eq, double K0, double DeltaK);
    double scalar_output(const imat &input_spins);
    void apply_filter(mat &input);
    mat calc_gradient_scalar_output(double h, const imat &input_spins);
    void update_weights(double eta, const mat &gradient);
};

#endif
