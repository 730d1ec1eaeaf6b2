// Host-side mirror of SPH::Particle (Source/Particle.h:50-265): the public fields drivers and GeneralAfter callbacks read
// and write, with the defaults of the reference constructor (Source/Particle.cpp:28-158). The authoritative state lives
// on the device in structure-of-arrays form; Domain copies these mirrors up before Solve and refreshes them around the
// callbacks and at output times. ~0.5 kB per particle instead of the reference's 2372 B.
#ifndef WF_HOST_PARTICLE_H
#define WF_HOST_PARTICLE_H
#include "matvec.h"
#include "Material.h"

#define TH_BC_NONE        0
#define TH_BC_CONVECTION  1
#define TH_BC_CONDUCTION  2

namespace SPH {

class Particle {
public:
    Particle(int Tag, Vec3_t const & x0, Vec3_t const & v0, double Mass0, double Density0, double h0, bool Fixed = false)
    {
        ID = Tag; ID_orig = Tag; x = x0; v = v0; Mass = Mass0; Density = Density0; RefDensity = Density0; h = h0; IsFree = !Fixed;
    }
    // ---- identity / flags
    int ID = 0, ID_orig = 0, Material = 0, Material_model = BILINEAR, Thermal_BC = TH_BC_NONE, Nb = 0, ContNb = 0, mesh = -1, element = -1;
    int inner_mirr_part = -1, ghost_plane_axis = 2;       // Particle.h:223-225: the inner particle a ghost mirrors, and the mirror plane's axis
    bool IsFree = true, Shepard = false, NoSlip = false, not_write_surf_ID = false, is_ghost = false, is_fixed = false, print_history = false;
    size_t PresEq = 0, Fail = 0;
    Material_ * mat = nullptr;
    // ---- kinematics
    Vec3_t x, v, va, vb, a, VXSPH, Displacement, normal, contforce;
    // ---- state
    double Mass = 0., Density = 0., RefDensity = 0., dDensity = 0., Pressure = 0., h = 0.;
    double Cs = 0., P0 = 0., G = 0., K = 0., Sigmay = 0., Ep = 0., Alpha = 0., Beta = 0., TI = 0., TIn = 4.0, TIInitDist = 0.;
    double pl_strain = 0., delta_pl_strain = 0., eff_strain_rate = 0., Sigma_eq = 0.;
    Mat3_t Sigma, ShearStress, StrainRate, RotationRate, Strain, Strain_pl, TIR;
    // ---- thermal
    double T = 0., k_T = 0., cp_T = 0., dTdt = 0., q_source = 0., q_conv = 0., T_inf = 0., h_conv = 0., q_plheat = 0., th_exp = 0.;
    double delta_cont = 0.;
    // ---- members some drivers assign that nothing on this path reads (Particle.h:80,83,167): v_max, hfac (smoothing-length update, never
    // called by the new solvers), Et_m (written by CalcStressStrain, read only in commented-out contact code, Contact.cpp:415)
    Vec3_t v_max; double hfac = 0., Et_m = 0., Et = 0.;
    bool wf_pooled = false;       // allocated from Domain's particle arena (contiguous in creation order) instead of the heap
    double mcp_t = 1.;            // Particle.cpp:149: mass * cp of a rigid-mesh particle; only read where tot_cont_heat_cond is non-zero, which CalcContactForcesWang never makes it (Contact_Wang.cpp:327,347)
};

} // namespace SPH
#endif
