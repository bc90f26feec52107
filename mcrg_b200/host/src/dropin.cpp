// Implementation of the drop-in classes (mcrg_dropin.hpp) over the C ABI of libmcrg_b200.so.
// Host code only orchestrates: every lattice operation of the hot path is a call into the device library, and a
// failing call throws — there is no CPU implementation of the path here.
#include "mcrg_dropin.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <stdexcept>

#include "../../../include/mcrg_b200.h"

// ---------------------------------------------------------------------------------------------------------
// definitions.cpp
// ---------------------------------------------------------------------------------------------------------

std::random_device rd;
std::mt19937_64 rng(rd());  // definitions.cpp:3-4: non-deterministic seed, as in the reference
std::uniform_int_distribution<int> binary(0, 1);
std::uniform_real_distribution<double> unif(0.0, 1.0);
std::normal_distribution<double> norm(0.0, 1.0);

vec flatten(mat M) {
    // definitions.cpp:9-19: columns laid end to end
    const int n = (int)M.cols();
    vec v(n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) v(j * n + i) = M(i, j);
    return v;
}

mat unflatten(vec v) {
    // definitions.cpp:21-31
    const int n = (int)std::floor(std::sqrt((double)v.size()));
    mat M(n, n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) M(i, j) = v(j * n + i);
    return M;
}

void display_spin_up() { printf("\033[34m%2s\033[0m", "O"); }
void display_spin_down() { printf("\033[37m%2s\033[0m", "X"); }

bool write_iter(int i) {
    // definitions.cpp:44-68: every iteration below 10, then every 10^k below 10^(k+1), then every 10^5
    int step = 1;
    for (int bound = 10; bound <= 100000; bound *= 10) {
        if (i < bound) return i % step == 0;
        step = bound;
    }
    return i % 100000 == 0;
}

std::string get_rounded_str(double num, int precision) {
    std::stringstream ss;
    ss << std::setprecision(precision) << num;
    return ss.str();
}

int split_samples(int rank, int n_processes, int n_samples) {
    // definitions.cpp:79-87 (including its behaviour when n_samples < n_processes^2: rank 0 takes the remainder)
    int n_loc = (int)std::ceil((double)n_samples / (double)n_processes);
    if (rank == 0) n_loc = n_samples - (n_processes - 1) * n_loc;
    return n_loc;
}

// ---------------------------------------------------------------------------------------------------------
// device plumbing
// ---------------------------------------------------------------------------------------------------------

namespace mcrg_b200 {

static void ck(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + mcrg_last_error());
}

static long env_long(const char *name, long dflt) {
    const char *e = std::getenv(name);
    return (e && *e) ? std::atol(e) : dflt;
}

Settings &settings() {
    static Settings s = {(int)env_long("MCRG_REPLICAS", 1024), (int)env_long("MCRG_SWEEPS_PER_UPDATE", 1),
                         (int)env_long("MCRG_DEVICE", 0), (std::uint64_t)env_long("MCRG_SEED", 12345),
                         (int)env_long("MCRG_QUIET", 0), -1, 1, 1};
    static bool parsed = false;
    if (!parsed) {
        s.devices = (int)env_long("MCRG_DEVICES", 1);
        if (s.devices < 1) s.devices = 1;
        // MCRG_UPDATE: "cluster" | "metropolis" | unset = auto (cluster updates — the reference's update family,
        // ising.cpp:87-155 — for N >= 32, where one Metropolis sweep is far from one Wolff update: tau ~ L^2.17 sweeps;
        // Metropolis sweeps below, where the resident kernel decorrelates a lattice in a few sweeps)
        const char *u = std::getenv("MCRG_UPDATE");
        s.cluster = -1;
        if (u && (std::string(u) == "cluster" || std::string(u) == "sw")) s.cluster = 1;
        if (u && std::string(u) == "metropolis") s.cluster = 0;
        s.compat = (int)env_long("MCRG_COMPAT", 1);
        parsed = true;
    }
    if (s.replicas < 1) s.replicas = 1;
    if (s.sweeps_per_update < 1) s.sweeps_per_update = 1;
    return s;
}

struct DeviceBatch {
    mcrg_ctx *ctx = nullptr;
    int L = 0, replicas = 0;
    DeviceBatch(int L_, int replicas_, std::uint32_t replica_base, int device = -1) : L(L_), replicas(replicas_) {
        ck(mcrg_ctx_create(device < 0 ? settings().device : device, L, replicas, settings().seed, replica_base, 1, &ctx), "mcrg_ctx_create");
        if (use_cluster(L)) ck(mcrg_set_update(ctx, MCRG_UPDATE_CLUSTER), "mcrg_set_update");
    }
    static bool use_cluster(int L) { return settings().cluster == 1 || (settings().cluster < 0 && L >= 32); }
    ~DeviceBatch() { mcrg_ctx_destroy(ctx); }
    DeviceBatch(const DeviceBatch &) = delete;
    DeviceBatch &operator=(const DeviceBatch &) = delete;
};

// Metropolis chosen explicitly on a large lattice: one sweep is not one Wolff update there (tau ~ N^2.17 sweeps at K_c), so an
// equilibration of `n_updates` updates from a hot start may leave the chains unequilibrated — say so once per call
static void warn_if_metropolis_is_short(int N, long n_updates) {
    if (DeviceBatch::use_cluster(N) || settings().quiet || N < 64) return;
    const double tau = std::pow((double)N, 2.17);
    if ((double)n_updates * settings().sweeps_per_update < tau)
        fprintf(stderr, "mcrg_b200: note: %ld Metropolis update(s) x %d sweep(s) on N = %d is short of the relaxation time (~N^2.17 = %.0f sweeps at K_c); "
                        "MCRG_UPDATE=cluster (the default for N >= 32) or a larger MCRG_SWEEPS_PER_UPDATE avoids biased, autocorrelated samples\n",
                n_updates, settings().sweeps_per_update, N, tau);
}

// Lattice objects get Philox replica ids above the range the drivers use for their batches
static std::atomic<std::uint32_t> g_next_lattice_id{0x40000000u};
// successive driver calls must not reuse Philox streams: each call takes a fresh block of replica ids
static std::atomic<std::uint32_t> g_next_batch_base{0};

static std::uint32_t take_batch_base(int replicas) { return g_next_batch_base.fetch_add((std::uint32_t)replicas); }

// exact 128-bit accumulator -> long double (64-bit mantissa: ample for covariances of ~1e-3 relative size)
static long double to_ld(std::int64_t hi, std::uint64_t lo) { return (long double)hi * 18446744073709551616.0L + (long double)lo; }

typedef __int128 i128;
static i128 to_i128(std::int64_t hi, std::uint64_t lo) { return ((i128)hi << 64) | (i128)lo; }
static long double to_ld(i128 x) { return to_ld((std::int64_t)(x >> 64), (std::uint64_t)x); }

struct Totals {
    std::vector<long double> v;  // [n_slots], summed over replicas (each one rounding of the exact total)
    std::vector<i128> exact;     // the same totals as exact integers
    std::vector<std::vector<long double>> per_replica;
};

static Totals fetch_totals(DeviceBatch &b) {
    mcrg_acc_layout lay;
    ck(mcrg_accumulators_layout(&lay), "mcrg_accumulators_layout");
    const size_t n = (size_t)b.replicas * lay.n_slots;
    std::vector<std::int64_t> hi(n);
    std::vector<std::uint64_t> lo(n);
    ck(mcrg_accumulators_get(b.ctx, hi.data(), lo.data()), "mcrg_accumulators_get");
    Totals t;
    t.exact.assign(lay.n_slots, 0);
    t.per_replica.assign(b.replicas, std::vector<long double>(lay.n_slots));
    for (int r = 0; r < b.replicas; ++r)
        for (int s = 0; s < lay.n_slots; ++s) {
            const i128 x = to_i128(hi[(size_t)r * lay.n_slots + s], lo[(size_t)r * lay.n_slots + s]);
            t.per_replica[r][s] = to_ld(x);
            t.exact[s] += x;  // totals are formed in exact integers: a running long-double sum rounds at every step once a
        }                     // slot exceeds 2^64 (N = 4096 after a few thousand samples, N = 16384 after ~32)
    t.v.resize(lay.n_slots);
    for (int s = 0; s < lay.n_slots; ++s) t.v[s] = to_ld(t.exact[s]);
    return t;
}

// The chains of one driver call spread over MCRG_DEVICES GPUs of this process (the role of `mpirun -n P` across nodes):
// consecutive blocks of replica ids, one context per device, all calls asynchronous so the devices work concurrently;
// the totals come from one NCCL all-reduce (mcrg_allreduce_accumulators = the MPI_Allreduce of mcrg.cpp:101-103), the
// per-chain sums (for the jackknife) from each device.
struct DeviceGroup {
    std::vector<std::unique_ptr<DeviceBatch>> parts;
    std::vector<mcrg_ctx *> ctxs;
    int replicas = 0;
    DeviceGroup(int L, int R, std::uint32_t base) : replicas(R) {
        int n_dev = settings().devices;
        if (n_dev > R) n_dev = R;
        int done = 0;
        for (int d = 0; d < n_dev; ++d) {
            const int r = (R - done) / (n_dev - d);
            parts.emplace_back(new DeviceBatch(L, r, base + (std::uint32_t)done, n_dev > 1 ? d : -1));
            ctxs.push_back(parts.back()->ctx);
            done += r;
        }
        if (n_dev > 1) ck(mcrg_comm_init_all(n_dev, ctxs.data()), "mcrg_comm_init_all");
    }
    template <typename F>
    void each(F f) {
        for (auto &p : parts) f(*p);
    }
    Totals totals() {
        Totals t;
        for (auto &p : parts) {
            Totals part = fetch_totals(*p);
            if (t.exact.empty()) t.exact.assign(part.exact.size(), 0);
            for (size_t s = 0; s < part.exact.size(); ++s) t.exact[s] += part.exact[s];
            for (auto &row : part.per_replica) t.per_replica.push_back(std::move(row));
        }
        t.v.resize(t.exact.size());
        for (size_t s = 0; s < t.exact.size(); ++s) t.v[s] = to_ld(t.exact[s]);
        if (parts.size() > 1) {  // the exact grand totals, reduced on the devices
            mcrg_acc_layout lay;
            ck(mcrg_accumulators_layout(&lay), "mcrg_accumulators_layout");
            std::vector<std::int64_t> hi(lay.n_slots);
            std::vector<std::uint64_t> lo(lay.n_slots);
            ck(mcrg_allreduce_accumulators((int)ctxs.size(), ctxs.data(), hi.data(), lo.data()), "mcrg_allreduce_accumulators");
            for (int s = 0; s < lay.n_slots; ++s)  // exact integers on both sides: the collective is exact and order independent
                if (to_i128(hi[s], lo[s]) != t.exact[s]) throw std::runtime_error("all-reduced totals differ from the sum of the per-chain sums");
        }
        return t;
    }
};

// mcrg.cpp:111-131 for the reference's operator pair (NN, NNN): A = <SbSb>-<Sb><Sb>^T, B = <SbS>-<Sb><S>^T,
// T = A^-1 B, lambda = larger real part of T's eigenvalues
static std::vector<double> lambdas_from(const std::vector<long double> &v, int n_lv) {
    mcrg_acc_layout lay;
    mcrg_accumulators_layout(&lay);
    const long double n = v[lay.slot_n];
    std::vector<double> out(n_lv, NAN);
    for (int lv = 0; lv < n_lv; ++lv) {
        long double A[2][2], B[2][2];
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                const long double sb_a = v[lay.slot_s + (lv + 1) * MCRG_NOP + a] / n, sb_b = v[lay.slot_s + (lv + 1) * MCRG_NOP + b] / n;
                const long double s_b = v[lay.slot_s + lv * MCRG_NOP + b] / n;
                A[a][b] = v[lay.slot_ss + (lv + 1) * 9 + b * MCRG_NOP + a] / n - sb_a * sb_b;
                B[a][b] = v[lay.slot_sbs + lv * 9 + b * MCRG_NOP + a] / n - sb_a * s_b;
            }
        const long double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
        if (det == 0.0L) continue;
        const long double Ai[2][2] = {{A[1][1] / det, -A[0][1] / det}, {-A[1][0] / det, A[0][0] / det}};
        long double T[2][2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) T[i][j] = Ai[i][0] * B[0][j] + Ai[i][1] * B[1][j];
        const long double tr = T[0][0] + T[1][1], dt = T[0][0] * T[1][1] - T[0][1] * T[1][0];
        const long double disc = tr * tr / 4 - dt;
        out[lv] = (double)(disc < 0 ? tr / 2 : tr / 2 + std::sqrt(disc));
    }
    return out;
}

}  // namespace mcrg_b200

using mcrg_b200::ck;
using mcrg_b200::DeviceBatch;
using mcrg_b200::settings;

// ---------------------------------------------------------------------------------------------------------
// Lattice  (lattice.cpp)
// ---------------------------------------------------------------------------------------------------------

Lattice::Lattice(int N) {
    MPI_Comm_size(MPI_COMM_WORLD, &n_processes_);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank_);
    a_ = 1;
    N_ = N;
    n_spins_ = N * N;
    pick_site_ = std::uniform_int_distribution<int>(0, n_spins_ - 1);
    // hot start, lattice.cpp:33-41: i.i.d. fair spins from the host generator, i outer / j inner
    spins_.resize(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) spins_(i, j) = rand_spin();
}

Lattice::Lattice(int a, imat spins) {
    MPI_Comm_size(MPI_COMM_WORLD, &n_processes_);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank_);
    spins_ = spins;
    a_ = a;
    N_ = (int)spins_.rows();
    n_spins_ = N_ * N_;
    pick_site_ = std::uniform_int_distribution<int>(0, n_spins_ - 1);
}

Lattice::~Lattice() {}

std::shared_ptr<DeviceBatch> Lattice::device_batch() {
    if (!dev_) dev_ = std::make_shared<DeviceBatch>(N_, 1, mcrg_b200::g_next_lattice_id.fetch_add(1));
    return dev_;
}

void Lattice::display_spins() {
    if (rank_ != 0) return;
    for (int i = 0; i < N_; ++i) {
        for (int j = 0; j < N_; ++j) (spins_(i, j) == 1) ? display_spin_up() : display_spin_down();
        printf("\n");
    }
    printf("(Lattice spacing a = %i)\n\n", a_);
}

void Lattice::write_spins(FILE *fptr) {
    if (rank_ != 0) return;
    fprintf(fptr, "\n");
    for (int i = 0; i < N_; ++i) {
        fprintf(fptr, "# ");
        for (int j = 0; j < N_; ++j) fprintf(fptr, "%i, ", spins_(i, j));
        fprintf(fptr, "\n");
    }
}

int Lattice::choose_random_spin() { return pick_site_(rng); }

static void observe(Lattice &lat, std::int64_t out[4]) {
    auto b = lat.device_batch();
    ck(mcrg_set_spins_i32_colmajor(b->ctx, 0, 1, lat.spins_.data()), "mcrg_set_spins_i32_colmajor");
    ck(mcrg_observables(b->ctx, &out[0], &out[1], &out[2], &out[3]), "mcrg_observables");
}

double Lattice::calc_nearest_neighbor_interaction() {
    std::int64_t o[4];
    observe(*this, o);
    return (double)o[0];
}

vec2D Lattice::calc_interactions() {
    std::int64_t o[4];
    observe(*this, o);
    vec2D S;
    S(0) = (double)o[0];
    S(1) = (double)o[1];
    return S;
}

double Lattice::calc_plaquette_interaction() {
    std::int64_t o[4];
    observe(*this, o);
    return (double)o[2];
}

long long Lattice::sum_spins() {
    std::int64_t o[4];
    observe(*this, o);
    return (long long)o[3];
}

static int wrap(int x, int N) { return ((x % N) + N) % N; }

imat Lattice::nearest_neighbors(int i, int j) {
    // lattice.cpp:124-136: rows = (+i, -i, +j, -j), columns = (row index, column index)
    imat nn(4, 2);
    nn(0, 0) = wrap(i + 1, N_); nn(0, 1) = j;
    nn(1, 0) = wrap(i - 1, N_); nn(1, 1) = j;
    nn(2, 0) = i;               nn(2, 1) = wrap(j + 1, N_);
    nn(3, 0) = i;               nn(3, 1) = wrap(j - 1, N_);
    return nn;
}

imat Lattice::next_nearest_neighbors(int i, int j) {
    // lattice.cpp:139-151
    imat nnn(4, 2);
    nnn(0, 0) = wrap(i + 1, N_); nnn(0, 1) = wrap(j + 1, N_);
    nnn(1, 0) = wrap(i - 1, N_); nnn(1, 1) = wrap(j + 1, N_);
    nnn(2, 0) = wrap(i + 1, N_); nnn(2, 1) = wrap(j - 1, N_);
    nnn(3, 0) = wrap(i - 1, N_); nnn(3, 1) = wrap(j - 1, N_);
    return nnn;
}

// ---------------------------------------------------------------------------------------------------------
// IsingModel  (ising.cpp)
// ---------------------------------------------------------------------------------------------------------

IsingModel::IsingModel(double K) {
    MPI_Comm_size(MPI_COMM_WORLD, &n_processes_);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank_);
    K_ = K;
    fptr_ = NULL;
}

static void device_sweeps(Lattice &lat, double K, int n_updates) {
    auto b = lat.device_batch();
    ck(mcrg_set_couplings(b->ctx, &K, 1), "mcrg_set_couplings");
    ck(mcrg_set_spins_i32_colmajor(b->ctx, 0, 1, lat.spins_.data()), "mcrg_set_spins_i32_colmajor");
    ck(mcrg_sweep(b->ctx, n_updates * settings().sweeps_per_update), "mcrg_sweep");
    ck(mcrg_get_spins_i32_colmajor(b->ctx, 0, 1, lat.spins_.data()), "mcrg_get_spins_i32_colmajor");
}

void IsingModel::sample_new_configuration(std::shared_ptr<Lattice> pLattice) { device_sweeps(*pLattice, K_, 1); }

double IsingModel::calc_energy(std::shared_ptr<Lattice> pLattice) {
    // ising.cpp:158-173 is K*S_nn/N^2 accumulated term by term; the integer S_nn is exact here
    return K_ * pLattice->calc_nearest_neighbor_interaction() / (pLattice->N_ * pLattice->N_);
}

double IsingModel::calc_magnetization(std::shared_ptr<Lattice> pLattice) {
    // ising.cpp:176-179: INTEGER division of the spin sum by N*N (so -1, 0 or +1)
    const int sum = (int)pLattice->sum_spins();
    return (double)(sum / (pLattice->N_ * pLattice->N_));
}

void IsingModel::equilibrate(std::shared_ptr<Lattice> pLattice, int n_samples_eq, bool write) {
    if (!write) {
        device_sweeps(*pLattice, K_, n_samples_eq);
        return;
    }
    // ising.cpp:22-74: thermodynamics log at log-spaced iterations, same file name and row format.  The reference averages
    // E, |M|, E^2, M^2 over its MPI ranks (MPI_Reduce, ising.cpp:50-53), every rank running its own chain; here the ranks
    // are MCRG_REPLICAS chains on the device: chain 0 is the caller's lattice, the others start hot like `Lattice(N)` on the
    // other ranks would (lattice.cpp:33-41).  Between two logged iterations the chains advance in one call; a logged
    // iteration is one level-0 measurement of all chains (exact integers S_nn, sum s), reduced on the host in the
    // reference's formulas.  MCRG_COMPAT=1 (default) keeps the reference's arithmetic, see Settings::compat.
    const int R = std::max(1, settings().replicas);
    const int N = pLattice->N_, spu = settings().sweeps_per_update;
    const bool compat = settings().compat != 0;
    DeviceBatch batch(N, R, mcrg_b200::take_batch_base(R));
    ck(mcrg_set_couplings(batch.ctx, &K_, 1), "mcrg_set_couplings");
    ck(mcrg_init_hot(batch.ctx), "mcrg_init_hot");
    ck(mcrg_set_spins_i32_colmajor(batch.ctx, 0, 1, pLattice->spins_.data()), "mcrg_set_spins_i32_colmajor");
    const std::string filename = "equilibrate_N_" + std::to_string(N) + "_K_" + get_rounded_str(K_, 7) + ".txt";
    fptr_ = fopen(filename.c_str(), "w");
    if (!fptr_) throw std::runtime_error("cannot open " + filename);
    fprintf(fptr_, "# Nearest neighbor coupling K = %lf\n", K_);
    fprintf(fptr_, "# Temperature T = %lf\n", -1 / K_);
    fprintf(fptr_, "# Number of lattice sites = %i\n", N * N);
    fprintf(fptr_, "# Lattice spacing = %i\n", pLattice->a_);
    fprintf(fptr_, "# Using %i parallel processes\n", R);
    fprintf(fptr_, "# %s, %s, %s, %s, %s, %s, %s\n", "Iteration", "Avg E/spin", "Stddev E/spin", "Heat Capacity",
            "Avg |M|/spin", "Stddev |M|/spin", "Susceptibility");
    std::vector<std::int64_t> S((size_t)R * 4);
    int done = 0;
    for (int n = 1; n <= n_samples_eq; ++n) {
        if (!write_iter(n) && n != n_samples_eq) continue;
        ck(mcrg_sweep(batch.ctx, (n - done) * spu), "mcrg_sweep");
        done = n;
        if (!write_iter(n)) break;
        ck(mcrg_measure(batch.ctx, 0, S.data(), nullptr), "mcrg_measure");
        double E_avg = 0, M_avg = 0, E2_avg = 0, M2_avg = 0;
        for (int r = 0; r < R; ++r) {
            const double E = K_ * (double)S[(size_t)r * 4 + 0] / ((double)N * N);                 // calc_energy, ising.cpp:158-173
            const long long sum = (long long)S[(size_t)r * 4 + 3];
            const double M = compat ? std::fabs((double)((int)sum / (N * N)))                   // ising.cpp:178: integer division
                                    : std::fabs((double)sum) / ((double)N * N);
            E_avg += E; M_avg += M; E2_avg += E * E; M2_avg += M * M;
        }
        E_avg /= R; M_avg /= R; E2_avg /= R; M2_avg /= R;                                       // ising.cpp:57-60
        double E_sigma = E2_avg - E_avg * E_avg, M_sigma = M2_avg - M_avg * M_avg;
        const double C = E_sigma * K_ * K_, Chi = -M_sigma * K_;
        E_sigma = std::sqrt(std::max(0.0, E_sigma));
        M_sigma = std::sqrt(std::max(0.0, M_sigma));
        const double per = compat ? (double)pLattice->n_spins_ : 1.0;                          // ising.cpp:72 divides again
        fprintf(fptr_, "%i, %10.7e, %10.7e, %10.7e, %10.7e, %10.7e, %10.7e\n", n, E_avg / per, E_sigma / per, C, M_avg / per,
                M_sigma / per, Chi);
    }
    ck(mcrg_get_spins_i32_colmajor(batch.ctx, 0, 1, pLattice->spins_.data()), "mcrg_get_spins_i32_colmajor");
    pLattice->write_spins(fptr_);
    fclose(fptr_);
    fptr_ = NULL;
}

// ---------------------------------------------------------------------------------------------------------
// MonteCarloRenormalizationGroup  (mcrg.cpp)
// ---------------------------------------------------------------------------------------------------------

MonteCarloRenormalizationGroup::MonteCarloRenormalizationGroup(int b) {
    b_ = b;
    fptr_ = NULL;
    iter_ = 0;
    rank_ = 0;
    n_processes_ = settings().replicas;  // independent chains on the device stand in for the reference's MPI ranks
    if (b != 2) throw std::invalid_argument("mcrg_b200: only the b = 2 block-spin transformation is implemented on the device");
    if (!settings().quiet) {
        printf("\n=============================================\n");
        printf("==========       MONTE CARLO       ==========\n");
        printf("==========  RENORMALIZATION GROUP  ==========\n");
        printf("=============================================\n\n");
        printf("* Using %i parallel process(es)\n", n_processes_);
        printf("* Scaling factor b = %i\n", b_);
    }
}

void MonteCarloRenormalizationGroup::calc_critical_exponent(int n_samples_eq, int n_samples, int N, double K) {
    const std::string filename = "critical_exponent_N_" + std::to_string(N) + "_K_" + get_rounded_str(K, 7) + ".txt";
    if (!settings().quiet) printf("* Calculating critical exponent at K = %lf\n\n", K);
    fptr_ = fopen(filename.c_str(), "w");
    if (!fptr_) throw std::runtime_error("cannot open " + filename);
    fprintf(fptr_, "# Number of parallel processes = %i\n", n_processes_);
    fprintf(fptr_, "# Number of equilibration samples = %i\n", n_samples_eq);
    fprintf(fptr_, "# Number of samples = %i\n", n_samples);
    fprintf(fptr_, "# %23s  %25s  %25s\n", "Blocking Level n", "Largest Eigenvalue", "Critical Exponent nu");

    const int R = n_processes_;
    const int per_replica = (n_samples + R - 1) / R;  // every chain takes ceil(n/R): SURVEY 7.0-9, no negative remainder
    const int n_lv = mcrg_levels_full(N);             // floor(log N / log b) - 1, mcrg.cpp:43
    const int spu = settings().sweeps_per_update;
    mcrg_b200::warn_if_metropolis_is_short(N, n_samples_eq);
    mcrg_b200::DeviceGroup group(N, R, mcrg_b200::take_batch_base(R));
    group.each([&](DeviceBatch &b) {
        ck(mcrg_set_couplings(b.ctx, &K, 1), "mcrg_set_couplings");
        ck(mcrg_init_hot(b.ctx), "mcrg_init_hot");                        // Lattice(N), mcrg.cpp:49
        ck(mcrg_sweep(b.ctx, n_samples_eq * spu), "mcrg_sweep");            // equilibrate, mcrg.cpp:50
    });
    if (!settings().quiet) printf("Sampling %i configurations...\n", n_samples);
    group.each([&](DeviceBatch &b) { ck(mcrg_run(b.ctx, per_replica, spu, n_lv, 0), "mcrg_run"); });  // mcrg.cpp:72-98
    mcrg_b200::Totals tot = group.totals();                                 // the all-reduce, mcrg.cpp:101-103

    lambdas_ = mcrg_b200::lambdas_from(tot.v, n_lv);
    nus_.assign(n_lv, NAN);
    lambda_errors_.assign(n_lv, NAN);
    // jackknife over chains (the reference prints point estimates only)
    if (R >= 8) {
        const int groups = std::min(R, 32);
        std::vector<std::vector<double>> loo;
        for (int g = 0; g < groups; ++g) {
            std::vector<long double> v = tot.v;
            for (int r = g; r < R; r += groups)
                for (size_t s = 0; s < v.size(); ++s) v[s] -= tot.per_replica[r][s];
            loo.push_back(mcrg_b200::lambdas_from(v, n_lv));
        }
        for (int lv = 0; lv < n_lv; ++lv) {
            double mean = 0, var = 0;
            for (auto &l : loo) mean += l[lv];
            mean /= groups;
            for (auto &l : loo) var += (l[lv] - mean) * (l[lv] - mean);
            lambda_errors_[lv] = std::sqrt(var * (groups - 1) / groups);
        }
    }
    double nu = 0.0;
    for (int n = 0; n < n_lv; ++n) {
        nu = std::log((double)b_) / std::log(lambdas_[n]);  // mcrg.cpp:131
        nus_[n] = nu;
        if (!settings().quiet) printf("n = %i: lambda = %lf, nu = %lf\n", n, lambdas_[n], nu);
        fprintf(fptr_, "%25i, %25.10lf, %25.10lf\n", n, lambdas_[n], nu);
    }
    if (!settings().quiet) printf("\n* Critical exponent: nu = %lf\n", nu);
    fclose(fptr_);
    fptr_ = NULL;
}

double MonteCarloRenormalizationGroup::locate_critical_point(int n_iterations, int n_samples_eq, int n_samples, int L, double K0) {
    if (!settings().quiet) printf("* Locating critical point starting from K0 = %lf\n", K0);
    const std::string filename = "critical_point_L_" + std::to_string(L) + "_K_" + get_rounded_str(K0, 7) + ".txt";
    fptr_ = fopen(filename.c_str(), "w");
    if (!fptr_) throw std::runtime_error("cannot open " + filename);
    fprintf(fptr_, "# Number of parallel processes = %i\n", n_processes_);
    fprintf(fptr_, "# Number of equilibration samples = %i\n", n_samples_eq);
    fprintf(fptr_, "# Number of samples = %i\n", n_samples);
    fprintf(fptr_, "# %23s  %25s  %25s  %25s\n", "Iteration", "Blocking Level n", "Starting K", "Approximate Kc");
    double K = K0;
    for (iter_ = 1; iter_ <= n_iterations; ++iter_) {
        if (!settings().quiet) printf("\nIteration %i:\n", iter_);
        K = approx_critical_point(n_samples_eq, n_samples, L, K);
    }
    if (!settings().quiet) {
        printf("\n* Critical point: Kc = %lf\n", K);
        printf("* Critical temperature: Tc = %lf\n", -1.0 / K);
    }
    fclose(fptr_);
    fptr_ = NULL;
    return K;
}

double MonteCarloRenormalizationGroup::approx_critical_point(int n_samples_eq, int n_samples, int L, double K) {
    // mcrg.cpp:191-310, Swendsen's two-lattice matching with the NN operator: lattice L blocked n+1 times is
    // compared with lattice L/b blocked n times.  Both batches run on the device; the six reductions of
    // mcrg.cpp:275-280 are reads of the accumulators.
    const int R = n_processes_;
    const int per_replica = (n_samples + R - 1) / R;
    const int nT = mcrg_levels_full(L);  // mcrg.cpp:196
    const int spu = settings().sweeps_per_update;
    const int S = L / b_;
    mcrg_acc_layout lay;
    ck(mcrg_accumulators_layout(&lay), "mcrg_accumulators_layout");

    mcrg_b200::warn_if_metropolis_is_short(L, n_samples_eq);
    DeviceBatch big(L, R, mcrg_b200::take_batch_base(R)), small(S, R, mcrg_b200::take_batch_base(R));
    for (DeviceBatch *b : {&big, &small}) {
        ck(mcrg_set_couplings(b->ctx, &K, 1), "mcrg_set_couplings");
        ck(mcrg_init_hot(b->ctx), "mcrg_init_hot");
        ck(mcrg_sweep(b->ctx, n_samples_eq * spu), "mcrg_sweep");
    }
    if (!settings().quiet) printf("Sampling %i configurations each...\n", n_samples);
    ck(mcrg_run(big.ctx, per_replica, spu, nT, 0), "mcrg_run");                       // levels 0..nT of L
    ck(mcrg_run(small.ctx, per_replica, spu, nT > 0 ? nT - 1 : 0, 0), "mcrg_run");    // levels 0..nT-1 of L/b
    const mcrg_b200::Totals tL = mcrg_b200::fetch_totals(big), tS = mcrg_b200::fetch_totals(small);
    // dK per blocking level from the totals of the two batches (mcrg.cpp:283-298)
    auto estimate = [&](const std::vector<long double> &vL, const std::vector<long double> &vS) {
        const long double n = vL[lay.slot_n];
        const long double SL_avg = vL[lay.slot_s + 0] / n, SS_avg = vS[lay.slot_s + 0] / n;
        std::vector<double> dK(nT);
        for (int k = 0; k < nT; ++k) {
            // SLb(k): NN sum of L blocked k+1 times; SSb(k): NN sum of L/b blocked k times (mcrg.cpp:253-263)
            const long double SLb = vL[lay.slot_s + (k + 1) * MCRG_NOP + 0] / n;
            const long double SSb = vS[lay.slot_s + k * MCRG_NOP + 0] / n;
            const long double SLb_SL = vL[lay.slot_sb0 + k * 9 + 0] / n;
            const long double SSb_SS = (k == 0 ? vS[lay.slot_ss + 0] : vS[lay.slot_sb0 + (k - 1) * 9 + 0]) / n;
            const long double dSL_dK = SLb_SL - SLb * SL_avg;  // mcrg.cpp:295
            const long double dSS_dK = SSb_SS - SSb * SS_avg;  // mcrg.cpp:296
            dK[k] = (double)((SLb - SSb) / (dSL_dK - dSS_dK));
        }
        return dK;
    };
    const std::vector<double> dK = estimate(tL.v, tS.v);
    // jackknife over groups of chains (chain r of the large batch and chain r of the small one leave together)
    kc_errors_.assign(nT, NAN);
    if (R >= 8) {
        const int groups = std::min(R, 32);
        std::vector<std::vector<double>> loo;
        for (int g = 0; g < groups; ++g) {
            std::vector<long double> vL = tL.v, vS = tS.v;
            for (int r = g; r < R; r += groups)
                for (size_t s = 0; s < vL.size(); ++s) {
                    vL[s] -= tL.per_replica[r][s];
                    vS[s] -= tS.per_replica[r][s];
                }
            loo.push_back(estimate(vL, vS));
        }
        for (int k = 0; k < nT; ++k) {
            double mean = 0, var = 0;
            for (auto &l : loo) mean += l[k];
            mean /= groups;
            for (auto &l : loo) var += (l[k] - mean) * (l[k] - mean);
            kc_errors_[k] = std::sqrt(var * (groups - 1) / groups);
        }
    }
    kcs_.assign(nT, NAN);
    double Kc = 0;
    for (int k = 0; k < nT; ++k) {
        Kc = K + dK[k];
        kcs_[k] = Kc;
        if (!settings().quiet) printf("n = %i: Kc = %lf\n", k, Kc);
        fprintf(fptr_, "%25i, %25i, %25.10lf, %25.10lf\n", iter_, k, K, Kc);
        fflush(fptr_);
    }
    return Kc;
}

std::shared_ptr<Lattice> MonteCarloRenormalizationGroup::block_spin_transformation(std::shared_ptr<Lattice> pLattice) {
    // mcrg.cpp:314-348 for b = 2: majority rule on the device, ties by Philox (keyed by the lattice's id and a
    // per-call counter so that repeated calls draw fresh coins)
    auto b = pLattice->device_batch();
    static std::atomic<std::uint64_t> calls{0};
    ck(mcrg_set_spins_i32_colmajor(b->ctx, 0, 1, pLattice->spins_.data()), "mcrg_set_spins_i32_colmajor");
    std::uint64_t t_saved = 0;  // the tie coins are keyed by the counter: use a private range, then put the lattice's own
    ck(mcrg_get_sweep_counter(b->ctx, &t_saved), "mcrg_get_sweep_counter");  // counter back so that its updates never reuse
    ck(mcrg_set_sweep_counter(b->ctx, (1ull << 40) + calls.fetch_add(1)), "mcrg_set_sweep_counter");  // Metropolis draws
    ck(mcrg_measure(b->ctx, 1, nullptr, nullptr), "mcrg_measure");
    const int Nb = pLattice->N_ / b_;
    imat block_spins(Nb, Nb);
    ck(mcrg_get_level_spins_i32_colmajor(b->ctx, 0, 1, block_spins.data()), "mcrg_get_level_spins_i32_colmajor");
    ck(mcrg_set_sweep_counter(b->ctx, t_saved), "mcrg_set_sweep_counter");
    return std::shared_ptr<Lattice>(new Lattice(pLattice->a_ * b_, block_spins));
}

// ---------------------------------------------------------------------------------------------------------
// RenormalizationGroupNeuralNetwork  (rgnn.cpp) — the consumer of the sampler; SURVEY 8f rank 2
// ---------------------------------------------------------------------------------------------------------

RenormalizationGroupNeuralNetwork::RenormalizationGroupNeuralNetwork(int b) {
    n_processes_ = settings().replicas;  // device chains stand in for the MPI ranks
    rank_ = 0;
    b_ = b;
    beta1_ = 0.9;
    beta2_ = 0.999;
    epsilon_ = 1E-8;
    eta_ = 0.0;
    w_ = 0.0;
    fptr_ = NULL;
    final_mse_ = 0.0;
    if (b != 2) throw std::invalid_argument("mcrg_b200: only b = 2 filters are implemented on the device");
    initialize();
}

void RenormalizationGroupNeuralNetwork::initialize() {
    // rgnn.cpp:19-35: N(0, 0.3^2) weights in (i outer, j inner) order, ADAM state cleared
    W_.resize(b_, b_);
    for (int i = 0; i < b_; ++i)
        for (int j = 0; j < b_; ++j) W_(i, j) = rand_weight();
    t_ = 1;
    m_.resize(b_, b_);
    v_.resize(b_, b_);
    m_.setZero();
    v_.setZero();
}

void RenormalizationGroupNeuralNetwork::set_weights(const mat &W) { W_ = W; }

void RenormalizationGroupNeuralNetwork::apply_filter(mat &input) {
    // rgnn.cpp:290-307: one filter application, each b x b block B -> entrywise L1 norm of the matrix product W*B.
    // Host helper for the "Example Flow" dump; the sampling loops evaluate the whole pyramid on the device.
    const int N = (int)input.rows() / b_;
    mat output(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double l1 = 0.0;
            for (int c = 0; c < b_; ++c)
                for (int r = 0; r < b_; ++r) {
                    double acc = 0.0;
                    for (int k = 0; k < b_; ++k) acc += W_(r, k) * input(b_ * i + k, b_ * j + c);
                    l1 += std::fabs(acc);
                }
            output(i, j) = l1;
        }
    input = output;
}

namespace {
struct RgnnEval {
    double u;
    double g[4];  // column-major 2x2
};
RgnnEval device_rgnn(const imat &spins, const mat &W, double h) {
    const int N = (int)spins.rows();
    DeviceBatch b(N, 1, mcrg_b200::g_next_lattice_id.fetch_add(1));
    ck(mcrg_set_spins_i32_colmajor(b.ctx, 0, 1, spins.data()), "mcrg_set_spins_i32_colmajor");
    ck(mcrg_rgnn_set_weights(b.ctx, W.data()), "mcrg_rgnn_set_weights");
    RgnnEval e;
    ck(mcrg_rgnn_eval(b.ctx, h, &e.u, e.g), "mcrg_rgnn_eval");
    return e;
}
}  // namespace

double RenormalizationGroupNeuralNetwork::scalar_output(const imat &input_spins) {
    return device_rgnn(input_spins, W_, 1e-4).u;  // rgnn.cpp:281-288
}

mat RenormalizationGroupNeuralNetwork::calc_gradient_scalar_output(double h, const imat &input_spins) {
    const RgnnEval e = device_rgnn(input_spins, W_, h);  // rgnn.cpp:310-339
    mat grad(b_, b_);
    for (int k = 0; k < 4; ++k) grad.data()[k] = e.g[k];
    return grad;
}

void RenormalizationGroupNeuralNetwork::update_weights(double eta, const mat &gradient) {
    // rgnn.cpp:342-357: ADAM with bias-corrected step size
    eta_ = eta * std::sqrt(1.0 - std::pow(beta2_, t_)) / (1.0 - std::pow(beta1_, t_));
    for (int i = 0; i < b_; ++i)
        for (int j = 0; j < b_; ++j) {
            m_(i, j) = beta1_ * m_(i, j) + (1.0 - beta1_) * gradient(i, j);
            v_(i, j) = beta2_ * v_(i, j) + (1.0 - beta2_) * gradient(i, j) * gradient(i, j);
            W_(i, j) -= eta_ * m_(i, j) / (std::sqrt(v_(i, j)) + epsilon_);
        }
    t_ += 1;
}

namespace {
// sums over all chains of one batch after n_loc samples each: {u, u^2, grad[4]}, chain order (deterministic)
void rgnn_block(DeviceBatch &b, const mat &W, int n_loc, int spu, double h, double out[6]) {
    ck(mcrg_rgnn_set_weights(b.ctx, W.data()), "mcrg_rgnn_set_weights");
    ck(mcrg_rgnn_accumulators_reset(b.ctx), "mcrg_rgnn_accumulators_reset");
    ck(mcrg_rgnn_run(b.ctx, n_loc, spu, h), "mcrg_rgnn_run");
    std::vector<double> s((size_t)b.replicas * 6);
    ck(mcrg_rgnn_accumulators_get(b.ctx, s.data()), "mcrg_rgnn_accumulators_get");
    for (int k = 0; k < 6; ++k) out[k] = 0.0;
    for (int r = 0; r < b.replicas; ++r)
        for (int k = 0; k < 6; ++k) out[k] += s[(size_t)r * 6 + k];
}

std::unique_ptr<DeviceBatch> equilibrated_batch(int L, int R, double K, int n_eq, int spu) {
    std::unique_ptr<DeviceBatch> b(new DeviceBatch(L, R, mcrg_b200::take_batch_base(R)));
    ck(mcrg_set_couplings(b->ctx, &K, 1), "mcrg_set_couplings");
    ck(mcrg_init_hot(b->ctx), "mcrg_init_hot");
    ck(mcrg_sweep(b->ctx, n_eq * spu), "mcrg_sweep");
    return b;
}
}  // namespace

void RenormalizationGroupNeuralNetwork::train_scalar_output(int L, int n_cycles, int n_samples, int n_samples_eq, double K,
                                                            double h, double eta) {
    // rgnn.cpp:46-190.  File name, header and row format as the reference writes them (rgnn.cpp:56-69, 154).
    const std::string filename = "train_scalar_b" + std::to_string(b_) + "_L" + std::to_string(L) + "_K" + get_rounded_str(K, 7) + ".txt";
    fptr_ = fopen(filename.c_str(), "w");
    if (!fptr_) throw std::runtime_error("cannot open " + filename);
    fprintf(fptr_, "# Initial Weights: ");
    for (int i = 0; i < b_; ++i)
        for (int j = 0; j < b_; ++j) fprintf(fptr_, "%20.10lf", W_(i, j));
    fprintf(fptr_, "\n# Cycles, Avg Output L, Var Output L, Avg Output S, Var Output S, MSE, || MSE Gradient ||\n");

    const int R = n_processes_, spu = settings().sweeps_per_update;
    const int n_loc = (n_samples + R - 1) / R;
    const double n_tot = (double)R * n_loc;
    auto big = equilibrated_batch(L, R, K, n_samples_eq, spu);         // rgnn.cpp:88-89
    auto small = equilibrated_batch(L / b_, R, K, n_samples_eq, spu);  // rgnn.cpp:92-94
    double mse = 0.0;
    for (int cycles = 0; cycles <= n_cycles; ++cycles) {
        double sL[6], sS[6];
        rgnn_block(*big, W_, n_loc, spu, h, sL);    // rgnn.cpp:106-130 for the large lattice ...
        rgnn_block(*small, W_, n_loc, spu, h, sS);  // ... and the small one
        const double uL_avg = sL[0] / n_tot, uS_avg = sS[0] / n_tot;
        const double uL_var = sL[1] / n_tot - uL_avg * uL_avg, uS_var = sS[1] / n_tot - uS_avg * uS_avg;
        mse = (uL_avg - uS_avg) * (uL_avg - uS_avg);  // rgnn.cpp:145
        mat grad(b_, b_);
        double gnorm = 0.0;
        for (int k = 0; k < 4; ++k) {
            grad.data()[k] = 2 * (uL_avg - uS_avg) * (sL[2 + k] / n_tot - sS[2 + k] / n_tot);  // rgnn.cpp:146
            gnorm += grad.data()[k] * grad.data()[k];
        }
        gnorm = std::sqrt(gnorm);
        if (cycles % 100 == 0 && !settings().quiet) printf("%10i%15.7e%15.7e%15.7e%15.7e\n", cycles, uL_avg, uS_avg, mse, gnorm);
        fprintf(fptr_, "%10i%15.7e%15.7e%15.7e%15.7e%15.7e%15.7e\n", cycles, uL_avg, uL_var, uS_avg, uS_var, mse, gnorm);
        update_weights(eta, grad);  // rgnn.cpp:159
    }
    final_mse_ = mse;
    fprintf(fptr_, "\n# Final Weights: ");
    for (int i = 0; i < b_; ++i)
        for (int j = 0; j < b_; ++j) fprintf(fptr_, "%20.10lf", W_(i, j));
    // rgnn.cpp:172-188: one more configuration of the large lattice pushed through the filter, level by level
    fprintf(fptr_, "\n# Example Flow: ");
    ck(mcrg_sweep(big->ctx, spu), "mcrg_sweep");
    imat spins(L, L);
    ck(mcrg_get_spins_i32_colmajor(big->ctx, 0, 1, spins.data()), "mcrg_get_spins_i32_colmajor");
    mat output = spins.cast<double>();
    while (output.rows() > 1) {
        for (int i = 0; i < output.rows(); ++i)
            for (int j = 0; j < output.cols(); ++j) fprintf(fptr_, "%20.10lf", output(i, j));
        fprintf(fptr_, ", ");
        apply_filter(output);
    }
    fprintf(fptr_, "%20.10lf", output(0, 0));
    fclose(fptr_);
    fptr_ = NULL;
}

void RenormalizationGroupNeuralNetwork::test_scalar_output(int L, int n_samples, int n_samples_eq, double K0, double DeltaK) {
    // rgnn.cpp:192-278
    const std::string filename = "test_scalar_b" + std::to_string(b_) + "_L" + std::to_string(L) + "_K" + get_rounded_str(K0, 7) + ".txt";
    fptr_ = fopen(filename.c_str(), "w");
    if (!fptr_) throw std::runtime_error("cannot open " + filename);
    fprintf(fptr_, "\n# Coupling K, Temperature T, Avg Output L, Var Output L, Avg Output S, Var Output S, MSE\n");
    const int R = n_processes_, spu = settings().sweeps_per_update;
    const int n_loc = (n_samples + R - 1) / R;
    const double n_tot = (double)R * n_loc;
    const double dK = DeltaK / 50;
    for (double K = K0 - DeltaK; K <= K0 + DeltaK; K += dK) {
        auto big = equilibrated_batch(L, R, K, n_samples_eq, spu);
        auto small = equilibrated_batch(L / b_, R, K, n_samples_eq, spu);
        double sL[6], sS[6];
        rgnn_block(*big, W_, n_loc, spu, 1e-4, sL);
        rgnn_block(*small, W_, n_loc, spu, 1e-4, sS);
        const double T = -1.0 / K;
        const double uL_avg = sL[0] / n_tot, uS_avg = sS[0] / n_tot;
        const double uL_var = sL[1] / n_tot - uL_avg * uL_avg, uS_var = sS[1] / n_tot - uS_avg * uS_avg;
        const double mse = (uL_avg - uS_avg) * (uL_avg - uS_avg);
        if (!settings().quiet) printf("%15.7e%15.7e%15.7e%15.7e%15.7e\n", K, T, uL_avg, uS_avg, mse);
        fprintf(fptr_, "%15.7e%15.7e%15.7e%15.7e%15.7e%15.7e%15.7e\n", K, T, uL_avg, uL_var, uS_avg, uS_var, mse);
        fflush(fptr_);
    }
    fclose(fptr_);
    fptr_ = NULL;
}
