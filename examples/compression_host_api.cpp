// Example / test driver for the host-side Domain API of weldform_b200 (weldform_b200/host/Domain.h).
//
// Written the way the reference's example drivers are (cf. examples/Compression_thermal.cpp in luchete80/WeldForm): build
// a cylinder with AddCylinderLength, assign per-particle material fields, install a GeneralAfter callback that drives
// the top rows and holds the bottom rows, call SolveDiffUpdateFraser. Everything under Solve runs on the GPU.
//
//   usage: compression_host_api <dx> <nsteps> <perturb_amp> <out.bin> [solver: 0 Fraser (default), 1 KickDrift, 2 LeapFrog,
//          3 ThermalSolve (heat conduction only, time step 0.05 s)] [ghost: 1 = AddCylinderLength(..., ghost = true), the bottom rows are
//          then left free and the mirror rows below the bottom plane hold the body (Domain::MoveGhost)] [pin_T: 1 = the callback also assigns T = 100 to the top rows]
// out.bin: int64 n, then n*(x[3], v[3], Displacement[3], pl_strain, T, Density, Sigma[9]) as float64 (creation order).
#include <cstdio>
#include "Domain.h"
#include "InteractionAlt.cpp"
#include "SolverFraser.cpp"
#include "SolverKickDrift.cpp"
#include "SolverLeapfrog.cpp"

#define TAU  0.005
#define VMAX 10.0

static bool pin_T = false;        // argv[7] = 1: the callback also holds the top rows at 100 degrees (a Dirichlet temperature through the bridge)

void UserAcc(SPH::Domain & domi)
{
    const double t = domi.getTime();
    const double f = t < TAU ? t / TAU : 1.0;
    for (size_t i = 0; i < domi.Particles.Size(); i++) {
        if (domi.Particles[i]->ID == 3) { domi.Particles[i]->a = Vec3_t(0.0, 0.0, 0.0); domi.Particles[i]->v = Vec3_t(0.0 * f, 0.0 * f, -VMAX * f); if (pin_T) domi.Particles[i]->T = 100.0; }
        if (domi.Particles[i]->ID == 2) { domi.Particles[i]->a = Vec3_t(0.0, 0.0, 0.0); domi.Particles[i]->v = Vec3_t(0.0, 0.0, 0.0); }
    }
}

int main(int argc, char ** argv) try
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s <dx> <nsteps> <perturb_amp> <out.bin>\n", argv[0]); return 2; }
    const double dx = std::atof(argv[1]); const int nsteps = std::atoi(argv[2]); const double amp = std::atof(argv[3]);
    SPH::Domain dom;
    dom.Dimension = 3; dom.Nproc = 4; dom.Kernel_Set(Qubic_Spline); dom.Scheme = 1;
    const double R = 0.15, L = 0.56, rho = 2700.0, K = 6.7549e10, G = 2.5902e10, Fy = 300.e6;
    const double h = dx * 1.2, Cs = sqrt(K / rho);
    dom.GeneralAfter = &UserAcc;
    dom.DomMax(0) = L; dom.DomMin(0) = -L;
    const bool ghost = argc > 6 && std::atoi(argv[6]) != 0;
    pin_T = argc > 7 && std::atoi(argv[7]) != 0;
    dom.AddCylinderLength(1, Vec3_t(0., 0., -L / 10.), R, L + 2. * L / 10., dx / 2., rho, h, false, ghost);
    for (size_t a = 0; a < dom.Particles.Size(); a++) {
        SPH::Particle * P = dom.Particles[a];
        P->G = G; P->PresEq = 0; P->k_T = 150.; P->cp_T = 960.; P->Cs = Cs; P->Shepard = false; P->Material = 2; P->Fail = 1; P->Sigmay = Fy;
        P->Alpha = 1.0; P->TI = 0.3; P->TIInitDist = dx; P->th_exp = 0.;
        const double x = P->x(0), y = P->x(1), z = P->x(2);
        P->T = 20.0 + 15.0 * sin(7.0 * x) * cos(5.0 * z);
        if (z < 0 && !ghost) P->ID = 2;
        if (z > L) P->ID = 3;
        if (amp != 0.0) P->v = Vec3_t(amp * sin(9.0 * y + 1) * cos(4.0 * z), amp * cos(8.0 * x - 1) * sin(3.0 * z), -amp * z * (1.0 + 0.3 * sin(11.0 * x * y)));
    }
    dom.m_kernel = SPH::iKernel(dom.Dimension, h);
    dom.BC.InOutFlow = 0;
    dom.thermal_solver = true; dom.pl_heating = true; dom.pl_work_heat_frac = 0.9; dom.auto_ts = false;
    const double timestep = 0.2 * h / Cs;
    const int solver = argc > 5 ? std::atoi(argv[5]) : 0;
    if (solver == 1) dom.SolveDiffUpdateKickDrift((nsteps - 0.5) * timestep, timestep, 1.e30, NULL, 1000000);
    else if (solver == 2) dom.SolveDiffUpdateLeapFrog((nsteps - 0.5) * timestep, timestep, 1.e30, NULL, 1000000);
    else if (solver == 3) dom.ThermalSolve((nsteps - 0.5) * 0.05, 0.05, 1.e30, NULL, 1000000);
    else dom.SolveDiffUpdateFraser((nsteps - 0.5) * timestep, timestep, 1.e30, NULL, 1000000);
    FILE * fp = std::fopen(argv[4], "wb");
    const long long n = (long long)dom.Particles.Size();
    std::fwrite(&n, sizeof(n), 1, fp);
    for (long long i = 0; i < n; ++i) {
        const SPH::Particle * P = dom.Particles[i];
        double rec[21]; int k = 0;
        for (int d = 0; d < 3; ++d) rec[k++] = P->x(d);
        for (int d = 0; d < 3; ++d) rec[k++] = P->v(d);
        for (int d = 0; d < 3; ++d) rec[k++] = P->Displacement(d);
        rec[k++] = P->pl_strain; rec[k++] = P->T; rec[k++] = P->Density;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) rec[k++] = P->Sigma(r, c);
        std::fwrite(rec, sizeof(double), 21, fp);
    }
    std::fclose(fp);
    std::printf("steps %zu wall %.3f s\n", dom.steps, dom.wall_seconds);
    return 0;
}
MECHSYS_CATCH
