// bench.py's end-to-end leg: the BASELINE workload (examples/Compression_thermal.cpp scaled by dx to N particles, plus the smooth synthetic
// velocity field of the device-resident leg so that the body yields within a few steps) driven through the reference-style C++ API --
// SPH::Domain::SolveDiffUpdateFraser of weldform_b200/host/Domain.h with a REAL GeneralAfter callback that loops over
// domi.Particles[i]->ID exactly like the reference's drivers do (examples/Compression_thermal.cpp:30-66), twice per step, on host Particle
// objects; the callback bridge (WELDFORM_SYNC = full | tagged | push) moves the particle data between host and device inside the timed region.
//
//   usage: bench_domain <particles> <steps> <warmup> [perturb_amp = 30] [callback threads = 16] [spin-up: 1 (default) wait until the body yields, 0 time from step 2] [callback: 0 (default) tests every particle's ID like the reference's examples, 1 visits index lists built at set-up]
// prints one line:  BENCH particles=<n> steps=<k> seconds=<s> first_step=<i> sync=<mode> max_pl=<..>
#include <cstdio>
#include <vector>
#include <omp.h>
#include "Domain.h"
#include "InteractionAlt.cpp"
#include "SolverFraser.cpp"

#define TAU  0.005
#define VMAX 10.0

// argv[7] = 1: the callback visits the rows it drives through index lists built once at set-up (what a driver written for ten million particles would
// do) instead of testing Particles[i]->ID of every particle like the reference's examples: shows how much of the e2e time is the walk itself
static std::vector<int> rows_top, rows_bottom;
void UserAccIndexed(SPH::Domain & domi)
{
    double vcompress;
    if (domi.getTime() < TAU) vcompress = VMAX / TAU * domi.getTime();
    else vcompress = VMAX;
    #pragma omp parallel for schedule (static) num_threads(domi.Nproc)
    for (size_t k = 0; k < rows_top.size(); k++) { SPH::Particle * P = domi.Particles[rows_top[k]]; P->a = Vec3_t(0.0, 0.0, 0.0); P->v = Vec3_t(0.0, 0.0, -vcompress); }
    #pragma omp parallel for schedule (static) num_threads(domi.Nproc)
    for (size_t k = 0; k < rows_bottom.size(); k++) { SPH::Particle * P = domi.Particles[rows_bottom[k]]; P->a = Vec3_t(0.0, 0.0, 0.0); P->v = Vec3_t(0.0, 0.0, 0.0); }
}

void UserAcc(SPH::Domain & domi)
{
    double vcompress;
    if (domi.getTime() < TAU) vcompress = VMAX / TAU * domi.getTime();
    else vcompress = VMAX;
    #pragma omp parallel for schedule (static) num_threads(domi.Nproc)
    for (size_t i = 0; i < domi.Particles.Size(); i++) {
        if (domi.Particles[i]->ID == 3) {
            domi.Particles[i]->a = Vec3_t(0.0, 0.0, 0.0);
            domi.Particles[i]->v = Vec3_t(0.0, 0.0, -vcompress);
        }
        if (domi.Particles[i]->ID == 2) {
            domi.Particles[i]->a = Vec3_t(0.0, 0.0, 0.0);
            domi.Particles[i]->v = Vec3_t(0.0, 0.0, 0.0);
        }
    }
}

int main(int argc, char ** argv) try
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s <particles> <steps> <warmup> [perturb_amp] [callback threads]\n", argv[0]); return 2; }
    const double n_target = std::atof(argv[1]); const size_t steps = (size_t)std::atoll(argv[2]), warmup = (size_t)std::atoll(argv[3]);
    const double amp = argc > 4 ? std::atof(argv[4]) : 30.0;
    SPH::Domain dom;
    dom.Dimension = 3; dom.Nproc = argc > 5 ? (size_t)std::atoi(argv[5]) : 16; dom.Kernel_Set(Qubic_Spline); dom.Scheme = 1;
    const double R = 0.15, L = 0.56, rho = 2700.0, K = 6.7549e10, G = 2.5902e10, Fy = 300.e6;
    const double dx = std::cbrt(M_PI * R * R * (L + 2.0 * L / 10.0) / n_target);
    const double h = dx * 1.2, Cs = sqrt(K / rho);
    dom.GeneralAfter = &UserAcc;
    dom.DomMax(0) = L; dom.DomMin(0) = -L;
    dom.AddCylinderLength(1, Vec3_t(0., 0., -L / 10.), R, L + 2. * L / 10., dx / 2., rho, h, false);
    for (size_t a = 0; a < dom.Particles.Size(); a++) {
        SPH::Particle * P = dom.Particles[a];
        P->G = G; P->PresEq = 0; P->k_T = 150.; P->cp_T = 960.; P->Cs = Cs; P->Shepard = false; P->Material = 2; P->Fail = 1; P->Sigmay = Fy;
        P->Alpha = 1.0; P->TI = 0.3; P->TIInitDist = dx; P->th_exp = 0.;
        const double x = P->x(0), y = P->x(1), z = P->x(2);
        P->T = 20.0 + 15.0 * sin(7.0 * x) * cos(5.0 * z);
        if (z < 0) { P->ID = 2; rows_bottom.push_back((int)a); }
        if (z > L) { P->ID = 3; rows_top.push_back((int)a); }
        if (amp != 0.0) P->v = Vec3_t(amp * sin(9.0 * y + 1) * cos(4.0 * z), amp * cos(8.0 * x - 1) * sin(3.0 * z), -amp * z * (1.0 + 0.3 * sin(11.0 * x * y)));
    }
    dom.m_kernel = SPH::iKernel(dom.Dimension, h);
    dom.BC.InOutFlow = 0;
    dom.thermal_solver = true; dom.pl_heating = true; dom.pl_work_heat_frac = 0.9; dom.auto_ts = false;
    const double timestep = 0.2 * h / Cs;
    if (argc > 7 && std::atoi(argv[7]) != 0) dom.GeneralAfter = &UserAccIndexed;
    dom.bench_steps = steps; dom.bench_warmup = warmup;
    if (argc > 6 && std::atoi(argv[6]) == 0) { dom.bench_spin_min = 2; dom.bench_wait_yield = false; }
    dom.SolveDiffUpdateFraser(1.0e30, timestep, 1.e30, NULL, 1000000);
    double max_pl = 0.0; for (size_t i = 0; i < dom.Particles.Size(); ++i) if (dom.Particles[i]->pl_strain > max_pl) max_pl = dom.Particles[i]->pl_strain;
    const char * m = std::getenv("WELDFORM_SYNC");
    std::printf("BENCH particles=%zu steps=%zu seconds=%.6f first_step=%zu sync=%s max_pl=%.17g\n", dom.Particles.Size(), steps, dom.bench_seconds, dom.bench_first_step, m ? m : "full", max_pl);
    return dom.bench_seconds > 0.0 ? 0 : 3;
}
MECHSYS_CATCH
