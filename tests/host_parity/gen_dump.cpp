// Test program, compiled TWICE by tests/test_host_api.py: against the reference's own headers (with oracle/shim, build container only) and
// against the host mirror (weldform_b200/host/). Prints, for each host-side lattice generator and material evaluation of the Domain API, every
// particle's position and mass (hex floats, exact) -- the two outputs must be identical.
#include <cstdio>
#include "Domain.h"

static void dump(const char * name, SPH::Domain & d, bool with_count = true)
{
    // Add3DCubicBoxParticles leaves solid_part_count at the constructor's -1 in the reference (its reductions then cover no particle at all,
    // Source/InteractionAlt.cpp:240); the mirror counts the particles -- the one deliberate difference, kept out of the comparison
    std::printf("@ %s n=%zu solid_part_count=%d\n", name, (size_t)d.Particles.Size(), with_count ? d.solid_part_count : 0);
    for (size_t i = 0; i < d.Particles.Size(); ++i) {
        const SPH::Particle * P = d.Particles[i];
        std::printf("@ %a %a %a %a %a %a %d %d\n", P->x(0), P->x(1), P->x(2), P->Mass, P->Density, P->h, P->ID, (int)P->IsFree);
    }
}

int main() try
{
    { SPH::Domain d; d.Dimension = 3; d.AddCylinderLength(1, Vec3_t(0., 0., -0.056), 0.15, 0.672, 0.02, 2700., 0.048, false); dump("AddCylinderLength", d); }
    { SPH::Domain d; d.Dimension = 3; d.AddCylinderLength(1, Vec3_t(0., 0., -0.056), 0.15, 0.672, 0.02, 2700., 0.048, false, true); dump("AddCylinderLength ghost", d);
      std::printf("@ GhostPairs %zu\n", (size_t)d.GhostPairs.Size());
      for (size_t g = 0; g < d.GhostPairs.Size(); ++g) std::printf("@ %zu %zu %d\n", (size_t)d.GhostPairs[g].first, (size_t)d.GhostPairs[g].second, (int)d.Particles[d.GhostPairs[g].second]->not_write_surf_ID); }
    for (int symz = 0; symz < 2; ++symz) {
      SPH::Domain d; d.Dimension = 3; d.AddXYSymCylinderLength(0, 0.0127, 0.015, 0.0006, 7850., 0.00144, false, symz != 0); dump(symz ? "AddXYSymCylinderLength symlength" : "AddXYSymCylinderLength", d);
      std::printf("@ GhostPairs %zu\n", (size_t)d.GhostPairs.Size());
      for (size_t g = 0; g < d.GhostPairs.Size(); ++g) { const SPH::Particle * P = d.Particles[d.GhostPairs[g].second];
          std::printf("@ %zu %zu %d %d %d\n", (size_t)d.GhostPairs[g].first, (size_t)d.GhostPairs[g].second, (int)P->not_write_surf_ID, P->ghost_plane_axis, P->inner_mirr_part); } }
    { SPH::Domain d; d.Dimension = 3; d.AddBoxLength(1, Vec3_t(-0.0155, -0.005, -0.005), 0.0331, 0.0101, 0.0101, 0.0005, 1000., 0.0013, 1, 0, false, false); dump("AddBoxLength", d); }
    { SPH::Domain d; d.Dimension = 3; d.Vol = 0.;
      d.Add3DCubicBoxParticles(1, Vec3_t(0., 0., 0.), 0.1, 0.05, 0.02, 0.005, 1000., 0.012);
      d.Add3DCubicBoxParticles(2, Vec3_t(0., 0., 0.02), 0.1, 0.05, 0.03, 0.005, 1000., 0.012);
      d.Calculate3DMass(1000.); dump("Add3DCubicBoxParticles x2 + Calculate3DMass", d, false); }
    { SPH::Domain d; d.Dimension = 3; d.AddTractionProbeLength(1, Vec3_t(0., 0., -0.0025), 0.00635, 0.012, 0.027, 0.0445, 0.0045, 0.0005, 7850., 0.0012, false); dump("AddTractionProbeLength", d); }
    { SPH::Domain d; d.Dimension = 3; d.AddCylUniformLength(0, 0.0127, 0.03, 0.001, 7850., 0.0024); dump("AddCylUniformLength", d); }
    { SPH::Domain d; d.Dimension = 3; d.AddCylUniformLength(0, 0.0127, 0.01, 0.001, 7850., 0.0024, M_PI / 8.0, 1, 0.00635); dump("AddCylUniformLength slice", d); }
    {   // material evaluations drivers use to seed Particle::Sigmay (Source/Material.cpp:6-73)
        Elastic_ el(70.e9, 0.3);
        Hollomon ho(el, 300.e6, 6.79e8, 0.15);
        JohnsonCook jc(el, 324.e6, 114.e6, 0.42, 0.002, 1.0, 1.34, 925., 294.);
        const double eps[] = { 0., 1e-4, 5e-3, 0.02, 0.3 };
        for (double e : eps) std::printf("@ hollomon %a %a\n", ho.CalcYieldStress(e), ho.CalcTangentModulus(e));
        for (double e : eps) for (double sr : { 0., 0.5, 10., 1.e4 }) for (double T : { 294., 400., 800. })
            std::printf("@ jc %a %a %a\n", jc.CalcYieldStress(e, sr, T), jc.CalcTangentModulus(e, sr, T), jc.CalcYieldStress(e));
        GMT gm(el, 1.0e-4, 0.15, 4.0e8, -0.002, 1.0e-5, 0.02, 1.0e-6, -0.002, 0.01, 3.0, 0.1, 1.0e5, 0.0, 500.0);          // Source/Material.cpp:96-161
        for (double e : eps) for (double sr : { 0., 0.5, 10., 1.e4 }) for (double T : { 20., 300., 800. })
            std::printf("@ gmt %a %a\n", gm.CalcYieldStress(e, sr, T), gm.CalcTangentModulus(e, sr, T));
        Material_ * base = &jc;
        std::printf("@ virtual %a %a\n", base->CalcYieldStress(0.01, 2.0, 350.), base->CalcTangentModulus(0.01, 2.0, 350.));
    }
    return 0;
}
catch (Fatal * f) { std::printf("@ Fatal\n"); return 1; }
