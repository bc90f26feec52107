// Command-line driver over the drop-in classes (the reference's drivers hard-code their parameters: main.cpp:8-13).
//   mcrg_app exponent N K n_eq n_samples            -> MonteCarloRenormalizationGroup::calc_critical_exponent
//   mcrg_app kc L K0 n_iterations n_eq n_samples    -> MonteCarloRenormalizationGroup::locate_critical_point
//   mcrg_app lattice N K n_updates                  -> Lattice / IsingModel fine-grained calls (prints observables)
//   mcrg_app train L K n_cycles n_samples n_eq      -> RenormalizationGroupNeuralNetwork::train_scalar_output
//   mcrg_app equilibrate N K n_eq                   -> IsingModel::equilibrate(lattice, n_eq, write = true): thermodynamics log
//   mcrg_app test L K0 DeltaK n_samples n_eq        -> RenormalizationGroupNeuralNetwork::test_scalar_output (W0 of train.cpp)
// Environment: MCRG_REPLICAS, MCRG_SWEEPS_PER_UPDATE, MCRG_SEED, MCRG_DEVICE, MCRG_QUIET.
#include <cstdlib>
#include <cstring>

#include "mcrg.hpp"
#include "rgnn.hpp"

int main(int argc, char **argv) {
    MPI_Init(NULL, NULL);
    if (argc < 2) {
        fprintf(stderr, "usage: %s exponent N K n_eq n_samples | kc L K0 n_it n_eq n_samples | lattice N K n_updates\n", argv[0]);
        return 2;
    }
    try {
        if (!strcmp(argv[1], "exponent") && argc == 6) {
            MonteCarloRenormalizationGroup rg(2);
            rg.calc_critical_exponent(atoi(argv[4]), atoi(argv[5]), atoi(argv[2]), atof(argv[3]));
            for (size_t n = 0; n < rg.lambdas_.size(); ++n)
                printf("RESULT level %zu lambda %.10f err %.10f nu %.10f\n", n, rg.lambdas_[n], rg.lambda_errors_[n], rg.nus_[n]);
        } else if (!strcmp(argv[1], "kc") && argc == 7) {
            MonteCarloRenormalizationGroup rg(2);
            const double Kc = rg.locate_critical_point(atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[2]), atof(argv[3]));
            printf("RESULT Kc %.10f\n", Kc);
            for (size_t n = 0; n < rg.kcs_.size(); ++n) printf("RESULT level %zu Kc %.10f err %.10f\n", n, rg.kcs_[n], rg.kc_errors_[n]);
        } else if (!strcmp(argv[1], "lattice") && argc == 5) {
            const int N = atoi(argv[2]);
            const double K = atof(argv[3]);
            std::shared_ptr<Lattice> lat(new Lattice(N));
            IsingModel ising(K);
            ising.equilibrate(lat, atoi(argv[4]), false);
            vec2D S = lat->calc_interactions();
            // recompute on the host from the public spins_ member with the reference's neighbour tables
            double snn = 0, snnn = 0;
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    imat nn = lat->nearest_neighbors(i, j), nnn = lat->next_nearest_neighbors(i, j);
                    for (int k = 0; k < 4; ++k) {
                        snn += lat->spins_(i, j) * lat->spins_(nn(k, 0), nn(k, 1));
                        snnn += lat->spins_(i, j) * lat->spins_(nnn(k, 0), nnn(k, 1));
                    }
                }
            printf("RESULT Snn %.0f %.0f Snnn %.0f %.0f E %.12f M %.0f sum %lld\n", S(0), snn, S(1), snnn, ising.calc_energy(lat),
                   ising.calc_magnetization(lat), lat->sum_spins());
        } else if (!strcmp(argv[1], "train") && argc == 7) {
            // mcrg_app train L K n_cycles n_samples n_eq  — train.cpp:19-31 with W0 = [.5 -.5; .5 -.5], h = 1e-4, eta = 1e-3
            mat W0(2, 2);
            W0(0, 0) = 0.5; W0(0, 1) = -0.5; W0(1, 0) = 0.5; W0(1, 1) = -0.5;
            RenormalizationGroupNeuralNetwork net(2);
            net.set_weights(W0);
            net.train_scalar_output(atoi(argv[2]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atof(argv[3]), 1e-4, 1e-3);
            printf("RESULT final_mse %.10e W %.10f %.10f %.10f %.10f\n", net.final_mse_, net.W_(0, 0), net.W_(0, 1), net.W_(1, 0), net.W_(1, 1));
        } else if (!strcmp(argv[1], "equilibrate") && argc == 5) {
            std::shared_ptr<Lattice> lat(new Lattice(atoi(argv[2])));
            IsingModel ising(atof(argv[3]));
            ising.equilibrate(lat, atoi(argv[4]), true);
            printf("RESULT sum %lld\n", lat->sum_spins());
        } else if (!strcmp(argv[1], "test") && argc == 7) {
            mat W0(2, 2);
            W0(0, 0) = 0.5; W0(0, 1) = -0.5; W0(1, 0) = 0.5; W0(1, 1) = -0.5;
            RenormalizationGroupNeuralNetwork net(2);
            net.set_weights(W0);
            net.test_scalar_output(atoi(argv[2]), atoi(argv[5]), atoi(argv[6]), atof(argv[3]), atof(argv[4]));
            printf("RESULT done\n");
        } else {
            fprintf(stderr, "bad arguments\n");
            return 2;
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "mcrg_app: %s\n", e.what());
        return 1;
    }
    MPI_Finalize();
    return 0;
}
