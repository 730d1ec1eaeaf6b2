// Per-particle phases of the Fraser step, contact, rigid-mesh update, reductions and the step orchestration
// (Source/SolverFraser.cpp:21-513). See include/weldform_b200.h for the reference member each entry point replaces.
#include <cub/cub.cuh>
#include <cstring>
#include <cmath>

#include "wf_internal.cuh"

using namespace wf;

namespace {

__device__ inline double ramp(double t, double tau) { if (tau <= 0.0) return 1.0; return t < tau ? t / tau : 1.0; }

struct StepArgs {
    int64_t n, cap; int nsph; double dt;
    double4 *Q0, *Q1, *Q2, *S0, *S1, *S2; double *F; int *I; const ClassProps *cls;
    const BCEntry *bcs; int n_bcs; DevScalars *sc;
};
#define OWN(f) a.F[(int64_t)(f) * a.cap + s]
#define OWNI(f) a.I[(int64_t)(f) * a.cap + s]

// device-side GeneralAfter (examples/Compression_thermal.cpp:30-66 pattern; oracle/ref_driver.cpp UserBC)
__device__ inline void apply_bc(const StepArgs &a, int64_t s, int id, double t, double &vx, double &vy, double &vz, double &ax, double &ay, double &az)
{
    for (int b = 0; b < a.n_bcs; ++b) {
        if (id == a.bcs[b].id) {
            const double f = ramp(t, a.bcs[b].tau);
            if (a.bcs[b].set_a) { ax = 0.0; ay = 0.0; az = 0.0; }
            vx = a.bcs[b].v[0] * f; vy = a.bcs[b].v[1] * f; vz = a.bcs[b].v[2] * f;
        }
    }
}

// mode bit 0: GeneralAfter before the move (SolverFraser.cpp:332); bit 1: position update (:337-343);
// bit 2: velocity update (:352-355); bit 3: GeneralAfter after the move (:360)
__global__ void k_bc_move(StepArgs a, int mode)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const double t = a.sc->time, dt = a.dt;
    double4 q = a.Q0[s], w = a.Q1[s];
    double ax = OWN(F_AX), ay = OWN(F_AY), az = OWN(F_AZ);
    const int id = OWNI(I_ID);
    if ((mode & 1) && a.n_bcs) apply_bc(a, s, id, t, w.x, w.y, w.z, ax, ay, az);
    if (mode & 2) {
        // du = (v + VXSPH)*deltat + 0.5*a*deltat*deltat with VXSPH == 0 (never computed by the new solvers)
        const double dux = (w.x + 0.0) * dt + 0.5 * ax * dt * dt, duy = (w.y + 0.0) * dt + 0.5 * ay * dt * dt, duz = (w.z + 0.0) * dt + 0.5 * az * dt * dt;
        OWN(F_DX) += dux; OWN(F_DY) += duy; OWN(F_DZ) += duz;
        q.x += dux; q.y += duy; q.z += duz;
        a.Q0[s] = q;
    }
    if (mode & 4) { w.x += ax * dt; w.y += ay * dt; w.z += az * dt; }
    if ((mode & 8) && a.n_bcs) apply_bc(a, s, id, t, w.x, w.y, w.z, ax, ay, az);
    a.Q1[s] = w;
    OWN(F_AX) = ax; OWN(F_AY) = ay; OWN(F_AZ) = az;
}

// Domain::MoveGhost (Domain.cpp:1142-1236), ghost_type Mirror_XYZ: the ghost particle takes v and a of its inner particle, mirrored across the
// plane normal to ghost_plane_axis, and its Sigma, Strain, ShearStress, StrainRate, RotationRate and Density. One thread per GhostPairs entry
// (creation indices -> slots through slot_of); inner particles are never ghosts, so the pairs are independent. The gathered caches that
// derive from the density (m/rho in Q1.w, P/rho^2 in Q2.w, with the ghost's own mass and pressure) follow the copy; Sigma/rho^2 (S0, S1)
// is rewritten by the ghost's own stress update before any momentum sum reads it.
// Several GPUs: a pair is handled by the rank that owns the GHOST; its inner particle must live on the same rank (the mirror rows and their partners
// sit within three rows of a plane, i.e. inside one slab unless the plane runs along the slab axis through a slab boundary). gp holds creation indices;
// out receives local indices (position in the rank's ascending creation-index order = the index space of slot_of), -1 where the particle is not local.
__global__ void k_ghost_localize(const int *gp, int64_t npairs, int64_t n_loc, const int *slot_of, const int *orig, int *out, int *flag)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= npairs) return;
    int lid[2];
    for (int m = 0; m < 2; ++m) {
        const int target = gp[3 * g + m];
        int64_t lo = 0, hi = n_loc;
        while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (orig[slot_of[mid]] < target) lo = mid + 1; else hi = mid; }
        lid[m] = (lo < n_loc && orig[slot_of[lo]] == target) ? (int)lo : -1;
    }
    out[3 * g] = lid[0]; out[3 * g + 1] = lid[1]; out[3 * g + 2] = gp[3 * g + 2];
    if (lid[1] >= 0 && lid[0] < 0) atomicExch(flag, 1);
}

__global__ void k_move_ghost(StepArgs a, const int *gp, int64_t npairs, const int *slot_of)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= npairs) return;
    if (gp[3 * g + 1] < 0) return;                   // several GPUs: the ghost of this pair lives on another rank
    const int64_t si = slot_of[gp[3 * g]], s = slot_of[gp[3 * g + 1]];
    const int axis = gp[3 * g + 2];
#define INNER(f) a.F[(int64_t)(f) * a.cap + si]
    const double4 wi = a.Q1[si];
    double v[3] = { wi.x, wi.y, wi.z }, acc[3] = { INNER(F_AX), INNER(F_AY), INNER(F_AZ) };
    v[axis] = -v[axis]; acc[axis] = -acc[axis];
    OWN(F_AX) = acc[0]; OWN(F_AY) = acc[1]; OWN(F_AZ) = acc[2];
    for (int k = 0; k < 6; ++k) { OWN(F_SG + k) = INNER(F_SG + k); OWN(F_ST + k) = INNER(F_ST + k); OWN(F_SH + k) = INNER(F_SH + k); OWN(F_SR + k) = INNER(F_SR + k); }
    for (int k = 0; k < 3; ++k) OWN(F_RR + k) = INNER(F_RR + k);
    double4 q2 = a.Q2[s];
    const double rho = a.Q2[si].y;
    q2.y = rho; q2.w = OWN(F_PRES) / (rho * rho);
    a.Q2[s] = q2;
    a.Q1[s] = make_double4(v[0], v[1], v[2], q2.x / rho);
#undef INNER
}

// Velocity kicks and the drift of the KickDrift / LeapFrog step orders, every particle (SolverKickDrift.cpp:225-259, SolverLeapfrog.cpp:241-286):
//   kind 0  v -= a*0.5*dt             (LeapFrog, first step only)      kind 1  v += a*0.5*prev_dt*factor   (KickDrift, factor = 1.)
//   kind 2  v += a*0.5*(prev_dt+dt)   (LeapFrog)                       kind 3  v += a*dt/2.*factor          (KickDrift)
//   kind 4  x_prev = x; du = (v + VXSPH)*dt*factor; Displacement += du; x += du   (both; VXSPH == 0, factor == 1.)
__global__ void k_kick(StepArgs a, int kind, double prev_dt, double dt)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    double4 w = a.Q1[s];
    if (kind == 4) {
        double4 q = a.Q0[s];
        const double dux = (w.x + 0.0) * dt * 1., duy = (w.y + 0.0) * dt * 1., duz = (w.z + 0.0) * dt * 1.;
        OWN(F_DX) += dux; OWN(F_DY) += duy; OWN(F_DZ) += duz;
        q.x += dux; q.y += duy; q.z += duz;
        a.Q0[s] = q;
        return;
    }
    const double ax = OWN(F_AX), ay = OWN(F_AY), az = OWN(F_AZ);
    if (kind == 0)      { w.x -= ax * 0.5 * dt; w.y -= ay * 0.5 * dt; w.z -= az * 0.5 * dt; }
    else if (kind == 1) { w.x += ax * 0.5 * prev_dt * 1.; w.y += ay * 0.5 * prev_dt * 1.; w.z += az * 0.5 * prev_dt * 1.; }
    else if (kind == 2) { const double f = prev_dt + dt; w.x += ax * 0.5 * f; w.y += ay * 0.5 * f; w.z += az * 0.5 * f; }
    else                { w.x += ax * dt / 2. * 1.; w.y += ay * dt / 2. * 1.; w.z += az * dt / 2. * 1.; }
    a.Q1[s] = w;
}

// mesh velocity ramps set from the callback in the oracle driver (trimesh[m]->SetVel)
__global__ void k_mesh_ramp(double *motion, int n_mesh, const DevScalars *sc)
{
    int m = threadIdx.x; if (m >= n_mesh) return;
    double *mo = motion + 16 * m;
    if (mo[10] != 0.0) { const double f = ramp(sc->time, mo[9]); mo[0] = mo[6] * f; mo[1] = mo[7] * f; mo[2] = mo[8] * f; }
}

// ---- CalcContactForcesWang (Source/Contact_Wang.cpp:15-348), one thread per SPH particle with contact candidates ----
struct ContactArgs {
    int K; const int *cont, *cont_cnt; const double *node, *enormal, *ev, *pplane; const int *en, *emesh;
    double friction_sta, friction_dyn, dfac, contact_hc; int cont_heat_fric, cont_heat_cond;
    int friction_function; double friction_m, friction_b;
    const int *slot_of; int first_mesh_local;          // element e's particle: slot_of[first_mesh_local + e] (its temperature, Contact_Wang.cpp:325)
    int n_mesh; double *mf_part; int *mf_flag;      // Domain::m_contact_force (Contact_Wang.cpp:312): per-block sums [block][WF_MAX_MESH_FORCES][3] + "block had a hit"
};
__global__ void __launch_bounds__(128) k_contact(StepArgs a, ContactArgs k)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // resets (Contact_Wang.cpp:27-42)
    double cfx = 0., cfy = 0., cfz = 0., dcont = 0., qfric = 0., qcont = 0., csum = 0., creac = 0., cext = 0.;
    double mf[WF_MAX_MESH_FORCES][3]; int hit = 0;
    const int cnt = s < a.n ? k.cont_cnt[s] : 0;
    if (cnt > 0) {
        for (int m = 0; m < WF_MAX_MESH_FORCES; ++m) mf[m][0] = mf[m][1] = mf[m][2] = 0.;
        const double4 q = a.Q0[s], w = a.Q1[s], q2 = a.Q2[s];
        const V3 x = v3(q.x, q.y, q.z), v = v3(w.x, w.y, w.z);
        const double h = q.w, Mass = q2.x, dens = q2.y, dt = a.dt;
        V3 acc = v3(OWN(F_AX), OWN(F_AY), OWN(F_AZ));
        for (int c = 0; c < cnt; ++c) {
            const int e = k.cont[(int64_t)c * a.cap + s];
            const V3 n2 = v3(k.enormal[3 * e], k.enormal[3 * e + 1], k.enormal[3 * e + 2]);
            const V3 v2 = v3(k.ev[3 * e], k.ev[3 * e + 1], k.ev[3 * e + 2]);
            const V3 vr = v - v2;
            const double delta_ = -dot(n2, vr);
            const V3 x_pred = (x + dt * v) + (dt * (dt * acc)) / 2.0;
            const double dist = dot(n2, x_pred) - k.pplane[e];
            if (dist < h) {
                const V3 qj = x - dist * n2;
                bool inside = true; int i = 0;
                while (i < 3 && inside) {
                    int j = i + 1; if (j > 2) j = 0;
                    const int ni = k.en[3 * e + i], nj = k.en[3 * e + j];
                    const V3 Ni = v3(k.node[3 * ni], k.node[3 * ni + 1], k.node[3 * ni + 2]), Nj = v3(k.node[3 * nj], k.node[3 * nj + 1], k.node[3 * nj + 2]);
                    const double crit = dot(cross(Nj - Ni, qj - Ni), n2);
                    if (crit < 0.0) inside = false;
                    i++;
                }
                if (inside) {
                    const double delta = h - dist;
                    const double kij = 2.0 * Mass / (dt * dt);
                    const double omega = sqrt(kij / Mass);
                    const double psi_cont = 2. * Mass * omega * k.dfac;
                    V3 cf = (kij * delta - psi_cont * delta_) * n2;
                    dcont = delta;
                    acc = acc + cf / Mass;
                    const double dS2 = pow(Mass / dens, 2.0 / 3.0);
                    double fr_sta = k.friction_sta, fr_dyn = k.friction_dyn;
                    if (k.friction_function == 1) fr_sta = fr_dyn = k.friction_m * q2.z + k.friction_b;      // Linear in T (Contact_Wang.cpp:243-244)
                    if (fr_sta > 0.) {
                        const V3 du = (x_pred - x) - dt * v2;
                        const V3 delta_tg = du - dot(du, n2) * n2;
                        const V3 tgforce = kij * delta_tg;
                        if (norm(tgforce) < fr_sta * norm(cf)) {
                            cf = cf - tgforce / Mass;          // unit quirk kept (Contact_Wang.cpp:279)
                            acc = acc - tgforce / Mass;
                        } else {
                            const V3 tgforce_dyn = ((fr_dyn * norm(cf)) * tgforce) / norm(tgforce);
                            cf = cf - tgforce_dyn;
                            acc = acc - tgforce_dyn / Mass;
                            if (k.cont_heat_fric) { const double abs_fv = fabs(dot(tgforce_dyn, vr)); qfric = abs_fv * dens / Mass; }
                        }
                    }
                    cfx = cf.x; cfy = cf.y; cfz = cf.z;
                    { const int m = k.emesh[e]; if (m < WF_MAX_MESH_FORCES) { mf[m][0] += cf.x; mf[m][1] += cf.y; mf[m][2] += cf.z; hit = 1; } }
                    csum += norm(cf);
                    creac += dot(acc, n2) * Mass;
                    cext += dot(cf, v2);
                    if (k.cont_heat_cond) qcont = dens * k.contact_hc * dS2 * (a.Q2[k.slot_of[k.first_mesh_local + e]].z - q2.z) / Mass;   // Particles[P2]->T: the element particle's temperature (the driver's value; nothing on this path changes it)
                }
            }
        }
        OWN(F_AX) = acc.x; OWN(F_AY) = acc.y; OWN(F_AZ) = acc.z;
    }
    // m_contact_force[m] += contforce, summed block-wise in a fixed order (no atomics: the per-mesh force is the same bits on every run)
    const int any = __syncthreads_or(hit);
    if (k.mf_flag) {
        if (threadIdx.x == 0) k.mf_flag[blockIdx.x] = any;
        if (any) {
            typedef cub::BlockReduce<double, 128> BR; __shared__ typename BR::TempStorage tmp;
            for (int m = 0; m < k.n_mesh && m < WF_MAX_MESH_FORCES; ++m)
                for (int d = 0; d < 3; ++d) {
                    const double r = BR(tmp).Sum(hit ? mf[m][d] : 0.);
                    __syncthreads();
                    if (threadIdx.x == 0) k.mf_part[((int64_t)blockIdx.x * WF_MAX_MESH_FORCES + m) * 3 + d] = r;
                }
        }
    }
    if (s >= a.n) return;
    OWN(F_CFX) = cfx; OWN(F_CFY) = cfy; OWN(F_CFZ) = cfz; OWN(F_DCONT) = dcont; OWN(F_QFRIC) = qfric; OWN(F_QCONT) = qcont;
    OWN(F_CSUM) = csum; OWN(F_CREAC) = creac; OWN(F_CEXT) = cext;
}

// ---- TriMesh::Update (Source/Mesh.cpp:273-293) + Domain::UpdateContactParticles (Source/Contact.cpp:41-55) ----------
__global__ void k_mesh_nodes_one(int n_nodes, double *node, double *node_v, const double *mo, double dt)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x; if (n >= n_nodes) return;
    const V3 m_v = v3(mo[0], mo[1], mo[2]), m_w = v3(mo[3], mo[4], mo[5]);
    V3 p = v3(node[3 * n], node[3 * n + 1], node[3 * n + 2]);
    const V3 vr = cross(m_w, p);
    const V3 nv = m_v + vr;
    p = p + dt * nv;
    node_v[3 * n] = nv.x; node_v[3 * n + 1] = nv.y; node_v[3 * n + 2] = nv.z;
    node[3 * n] = p.x; node[3 * n + 1] = p.y; node[3 * n + 2] = p.z;
}
__global__ void k_mesh_elems(int n_elems, int nsph, const double *node, const double *node_v, const int *en, const int *nfar, double *enormal, double *ev, double *pplane,
                             const int *slot_of, double4 *Q1, double *F, int64_t cap)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x; if (e >= n_elems) return;
    const int n0 = en[3 * e], n1 = en[3 * e + 1], n2 = en[3 * e + 2];
    const V3 A = v3(node[3 * n0], node[3 * n0 + 1], node[3 * n0 + 2]), B = v3(node[3 * n1], node[3 * n1 + 1], node[3 * n1 + 2]), C = v3(node[3 * n2], node[3 * n2 + 1], node[3 * n2 + 2]);
    const V3 u = B - A, v = C - A, w = cross(u, v);
    const V3 nr = w / norm(w);                                        // CalcNormals (Mesh.cpp:235-243)
    const int nf = en[3 * e + nfar[e]];
    pplane[e] = dot(v3(node[3 * nf], node[3 * nf + 1], node[3 * nf + 2]), nr);   // UpdatePlaneCoeff
    enormal[3 * e] = nr.x; enormal[3 * e + 1] = nr.y; enormal[3 * e + 2] = nr.z;
    V3 vv = v3(0., 0., 0.);
    vv = vv + v3(node_v[3 * n0], node_v[3 * n0 + 1], node_v[3 * n0 + 2]); vv = vv + v3(node_v[3 * n1], node_v[3 * n1 + 1], node_v[3 * n1 + 2]); vv = vv + v3(node_v[3 * n2], node_v[3 * n2 + 1], node_v[3 * n2 + 2]);
    vv = vv / 3;
    ev[3 * e] = vv.x; ev[3 * e + 1] = vv.y; ev[3 * e + 2] = vv.z;
    // element particle: v = mean node velocity, a = 0, normal = element normal
    const int s = slot_of[nsph + e];            // nsph = local index of the first element particle
    double4 q1 = Q1[s]; q1.x = vv.x; q1.y = vv.y; q1.z = vv.z; Q1[s] = q1;
    F[(int64_t)F_AX * cap + s] = 0.; F[(int64_t)F_AY * cap + s] = 0.; F[(int64_t)F_AZ * cap + s] = 0.;
    F[(int64_t)F_NX * cap + s] = nr.x; F[(int64_t)F_NY * cap + s] = nr.y; F[(int64_t)F_NZ * cap + s] = nr.z;
}

// ---- reductions ---------------------------------------------------------------------------------------------------------
enum { R_DKIN = 0, R_DINT, R_PLW, R_FRIC, R_COND, R_CSUM, R_CREAC, R_CEXT, R_NSUM, R_MAXPL = R_NSUM, R_MD0, R_MD1, R_MD2, R_TMAX, R_TMIN, R_NTOT, R_MCF = R_NTOT };   // slots R_MCF .. R_MCF + 3*WF_MAX_MESH_FORCES - 1: m_contact_force
static_assert(R_MCF + 3 * WF_MAX_MESH_FORCES <= 32, "scalar record is 32 doubles per rank");

__global__ void __launch_bounds__(256) k_reduce_partial(StepArgs a, int first_fem, double *partial)
{
    double acc[R_NTOT];
    for (int k = 0; k < R_NSUM; ++k) acc[k] = 0.;
    acc[R_MAXPL] = 0.; acc[R_MD0] = acc[R_MD1] = acc[R_MD2] = 0.; acc[R_TMAX] = 0.; acc[R_TMIN] = 1000.;
#pragma unroll 2
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < a.n; s += (int64_t)gridDim.x * blockDim.x) {
        const int o = OWNI(I_ORIG);
        acc[R_DKIN] += OWN(F_DKIN);
        acc[R_MAXPL] = fmax(acc[R_MAXPL], OWN(F_PLS));
        if (o < a.nsph) {
            const double m = a.Q2[s].x, d = a.Q2[s].y;
            acc[R_PLW] += m / d * OWN(F_QPL);
            acc[R_FRIC] += OWN(F_QFRIC) * m / d;
            acc[R_COND] += OWN(F_QCONT) * m / d;
            acc[R_CSUM] += OWN(F_CSUM); acc[R_CREAC] += OWN(F_CREAC); acc[R_CEXT] += OWN(F_CEXT);
            const double dx = OWN(F_DX), dy = OWN(F_DY), dz = OWN(F_DZ);
            acc[R_MD0] = fmax(acc[R_MD0], dx * dx); acc[R_MD1] = fmax(acc[R_MD1], dy * dy); acc[R_MD2] = fmax(acc[R_MD2], dz * dz);
            const double T = a.Q2[s].z; acc[R_TMAX] = fmax(acc[R_TMAX], T); acc[R_TMIN] = fmin(acc[R_TMIN], T);
        }
        if (o < first_fem) acc[R_DINT] += OWN(F_DINT);
    }
    typedef cub::BlockReduce<double, 256> BR; __shared__ typename BR::TempStorage tmp;
    for (int k = 0; k < R_NTOT; ++k) {
        double r;
        if (k < R_NSUM) r = BR(tmp).Sum(acc[k]);
        else if (k == R_TMIN) r = BR(tmp).Reduce(acc[k], cub::Min());
        else r = BR(tmp).Reduce(acc[k], cub::Max());
        __syncthreads();
        if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * 32 + k] = r;
    }
}
// stage 2: block partials -> this rank's totals (32 doubles)
__global__ void __launch_bounds__(256) k_reduce_rank(const double *partial, int nblocks, double *out, const double *mf_part, const int *mf_flag, int mf_blocks, int n_mesh)
{
    typedef cub::BlockReduce<double, 256> BR; __shared__ typename BR::TempStorage tmp;
    for (int q = 0; q < 3 * WF_MAX_MESH_FORCES; ++q) {          // per-mesh contact force: block partials of k_contact, fixed order
        double v = 0.;
        if (q < 3 * n_mesh) for (int b = threadIdx.x; b < mf_blocks; b += blockDim.x) if (mf_flag[b]) v += mf_part[(int64_t)b * 3 * WF_MAX_MESH_FORCES + q];
        const double r = (q < 3 * n_mesh && mf_blocks > 0) ? BR(tmp).Sum(v) : 0.;
        __syncthreads();
        if (threadIdx.x == 0) out[R_MCF + q] = r;
    }
    for (int k = 0; k < R_NTOT; ++k) {
        double v = (k < R_NSUM) ? 0. : (k == R_TMIN ? 1000. : 0.);
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            const double p = partial[(int64_t)b * 32 + k];
            if (k < R_NSUM) v += p; else if (k == R_TMIN) v = fmin(v, p); else v = fmax(v, p);
        }
        double r;
        if (k < R_NSUM) r = BR(tmp).Sum(v); else if (k == R_TMIN) r = BR(tmp).Reduce(v, cub::Min()); else r = BR(tmp).Reduce(v, cub::Max());
        __syncthreads();
        if (threadIdx.x == 0) out[k] = r;
    }
}
// stage 3: totals of all ranks (rank order, identical on every rank) -> device scalars.
// accumulate = 1: end-of-step bookkeeping (energy sums += inc*dt etc.); 0: refresh maxima only
__global__ void k_reduce_final(const double *all, int nranks, DevScalars *sc, double dt, int accumulate)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double res[R_NTOT];
    for (int k = 0; k < R_NTOT; ++k) {
        double v = (k < R_NSUM) ? 0. : (k == R_TMIN ? 1000. : 0.);
        for (int r = 0; r < nranks; ++r) { const double p = all[r * 32 + k]; if (k < R_NSUM) v += p; else if (k == R_TMIN) v = fmin(v, p); else v = fmax(v, p); }
        res[k] = v;
    }
    {
        sc->max_pl_strain = res[R_MAXPL];
        // Domain::max_disp is evaluated at the START of a step (SolverFraser.cpp:151-159): keep the previous end-of-step
        // value so that wf_reduce can report what the reference holds once the step has completed
        if (accumulate) { sc->max_disp2_prev[0] = sc->max_disp2[0]; sc->max_disp2_prev[1] = sc->max_disp2[1]; sc->max_disp2_prev[2] = sc->max_disp2[2]; }
        sc->max_disp2[0] = res[R_MD0]; sc->max_disp2[1] = res[R_MD1]; sc->max_disp2[2] = res[R_MD2];
        sc->max_T = res[R_TMAX]; sc->min_T = res[R_TMIN];
        if (accumulate) {
            sc->kin_energy_sum += res[R_DKIN] * dt;             // Domain.cpp:1807
            sc->int_energy_sum += res[R_DINT] * dt;             // Domain.cpp:1828
            sc->plastic_work += res[R_PLW] * dt;                // ThermalStuct.cpp:36
            sc->contact_friction_work += res[R_FRIC] * dt;      // Thermal.cpp:208
            sc->accum_cont_heat_cond += res[R_COND] * dt;       // Thermal.cpp:209
            sc->contact_force_sum = res[R_CSUM]; sc->contact_reaction_sum = res[R_CREAC];
            for (int q = 0; q < 3 * WF_MAX_MESH_FORCES; ++q) { double v = 0.; for (int r = 0; r < nranks; ++r) v += all[r * 32 + R_MCF + q]; sc->mcf[q / 3][q % 3] = v; }
            sc->ext_forces_work += res[R_CEXT] * dt;            // Contact_Wang.cpp:344
            if (accumulate == 1) sc->time += dt;                // SolverFraser.cpp:384 (accumulate == 2: k_advance_time already did, on the compute stream)
        }
    }
}
__global__ void k_set_time(DevScalars *sc, double t, double dt) { sc->time = t; sc->dt = dt; }
__global__ void k_advance_time(DevScalars *sc, double dt) { sc->time += dt; }

// InitialChecks (Domain.cpp:2097-2113): Pressure = EOS(...) for every particle
__global__ void k_init_pressure(StepArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const double4 q2 = a.Q2[s]; const ClassProps cp = a.cls[OWNI(I_CLS)];
    if (OWNI(I_ORIG) >= a.nsph) { OWN(F_PRES) = eos(cp.PresEq, cp.Cs, cp.P0, 0.0, cp.RefDensity); return; }
    const double P = eos(cp.PresEq, cp.Cs, cp.P0, q2.y, cp.RefDensity);
    OWN(F_PRES) = P;
    ((double *)&a.Q2[s])[3] = P / (q2.y * q2.y);
}

StepArgs make_step_args(wf_ctx *c)
{
    StepArgs a; a.n = c->n_loc; a.cap = c->cap; a.nsph = (int)c->nsph; a.dt = c->dt;
    a.Q0 = c->Q0[c->cur]; a.Q1 = c->Q1[c->cur]; a.Q2 = c->Q2[c->cur]; a.S0 = c->S0[c->cur]; a.S1 = c->S1[c->cur]; a.S2 = c->S2[c->cur];
    a.F = c->F[c->cur]; a.I = c->I[c->cur]; a.cls = c->d_classes; a.bcs = c->d_bcs; a.n_bcs = c->n_bcs; a.sc = c->d_sc;
    return a;
}

} // namespace

namespace wf {

int contact_impl(wf_ctx *c)
{
    if (!c->prm.contact) return 0;
    if (!c->cont) return fail(c, WF_ERR_ARG, "wf_contact before the first neighbour build");
    ProfScope ps(c, PK_CONTACT);
    StepArgs a = make_step_args(c);
    ContactArgs k; k.K = c->cont_k; k.cont = c->cont; k.cont_cnt = c->cont_cnt; k.node = c->dm.node; k.enormal = c->dm.enormal; k.ev = c->dm.ev; k.pplane = c->dm.pplane;
    k.en = c->dm.en; k.emesh = c->dm.emesh; k.friction_sta = c->prm.friction_sta; k.friction_dyn = c->prm.friction_dyn; k.friction_function = c->prm.friction_function; k.friction_m = c->prm.friction_m; k.friction_b = c->prm.friction_b; k.dfac = c->prm.dfac; k.contact_hc = c->prm.contact_hc;
    k.cont_heat_fric = c->prm.cont_heat_fric; k.cont_heat_cond = c->prm.cont_heat_cond;
    const int nblk = grid_for(c->n_loc, 128);
    if (nblk > c->mf_cap) {
        cudaFree(c->d_mf_part); cudaFree(c->d_mf_flag); c->d_mf_part = nullptr; c->d_mf_flag = nullptr; c->mf_cap = 0;
        const int want = nblk + nblk / 8 + 64;
        WF_CUDA(c, cudaMalloc(&c->d_mf_part, (size_t)want * 3 * WF_MAX_MESH_FORCES * sizeof(double))); WF_CUDA(c, cudaMalloc(&c->d_mf_flag, (size_t)want * sizeof(int)));
        c->mf_cap = want;
    }
    c->mf_blocks = nblk;
    k.n_mesh = (int)c->meshes.size(); k.mf_part = c->d_mf_part; k.mf_flag = c->d_mf_flag;
    k.slot_of = c->slot_of; k.first_mesh_local = (int)(c->n_loc - c->nmeshp);
    k_contact<<<grid_for(c->n_loc, 128), 128, 0, c->stream>>>(a, k); WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}

extern "C" int wf_fraser_run(wf_ctx *c, int64_t nsteps);
static int bc_move(wf_ctx *c, int mode, int pk)
{
    ProfScope ps(c, pk);
    StepArgs a = make_step_args(c);
    if ((mode & 9) && !c->meshes.empty()) { k_mesh_ramp<<<1, 32, 0, c->stream>>>(c->dm.motion, (int)c->meshes.size(), c->d_sc); WF_LAUNCH_CHECK(c); c->launches++; }
    k_bc_move<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a, mode); WF_LAUNCH_CHECK(c); c->launches++;
    c->dens_valid = false;       // x / v changed after the last k_post (if any): its prepared density rate is stale
    return 0;
}
int apply_bcs_impl(wf_ctx *c) { if (c->n_bcs == 0 && c->meshes.empty()) return 0; return bc_move(c, 1, PK_BC); }
int move_pos_impl(wf_ctx *c) { return bc_move(c, 2, PK_MOVE); }
int move_vel_impl(wf_ctx *c) { return bc_move(c, 4, PK_MOVE); }
int move_ghost_impl(wf_ctx *c)
{
    if (c->n_ghost == 0) return 0;
    ProfScope ps(c, PK_MOVE);
    StepArgs a = make_step_args(c);
    const int *pairs = c->d_ghost;
    if (c->comm.nranks > 1) { if (!c->ghost_loc_valid) WF_CHECK(ghost_localize(c)); pairs = c->d_ghost_loc; }
    k_move_ghost<<<grid_for(c->n_ghost, 128), 128, 0, c->stream>>>(a, pairs, c->n_ghost, c->slot_of); WF_LAUNCH_CHECK(c); c->launches++;
    c->dens_valid = false;
    return 0;
}
int ghost_localize(wf_ctx *c)
{
    if (c->n_ghost == 0 || c->comm.nranks <= 1) return 0;
    if (!c->d_ghost_loc) WF_CUDA(c, cudaMalloc(&c->d_ghost_loc, (size_t)3 * c->n_ghost * sizeof(int)));
    WF_CUDA(c, cudaMemsetAsync(c->d_maxcnt + 3, 0, sizeof(int), c->stream));
    k_ghost_localize<<<grid_for(c->n_ghost, 128), 128, 0, c->stream>>>(c->d_ghost, c->n_ghost, c->n_loc, c->slot_of, c->Ip(I_ORIG), c->d_ghost_loc, c->d_maxcnt + 3); WF_LAUNCH_CHECK(c); c->launches++;
    int bad = 0; WF_CUDA(c, cudaMemcpyAsync(&bad, c->d_maxcnt + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); WF_CUDA(c, cudaStreamSynchronize(c->stream));
    // every rank takes the same decision (a rank that stopped alone would leave the others waiting in the next halo exchange)
    double lo = (double)bad, hi = (double)bad; WF_CHECK(comm_global_minmax(c, &lo, &hi));
    if (hi != 0.0) return fail(c, WF_ERR_UNSUPPORTED, "a ghost (symmetry) particle and the inner particle it mirrors live on different GPUs: the mirror plane crosses a slab boundary");
    c->ghost_loc_valid = true;
    return 0;
}

int reduce_impl(wf_ctx *c, bool accumulate)
{
    ProfScope ps(c, PK_REDUCE);
    StepArgs a = make_step_args(c);
    int nb = std::min<int64_t>(c->partial_blocks, (c->n_loc + 255) / 256); if (nb < 1) nb = 1;
    const int first_fem = c->meshes.empty() ? 2147483647 : (int)c->meshes[0].first_particle;
    const bool multi = c->comm.nranks > 1;
    k_reduce_partial<<<nb, 256, 0, c->stream>>>(a, first_fem, c->d_partial); WF_LAUNCH_CHECK(c);
    if (multi && accumulate && !(c->prm.tuning & 2048)) {
        // several GPUs, end of a step: the all-gather of the 32 scalars and the accumulation run on the COMMUNICATION stream; the compute
        // stream only advances the time (which the boundary-condition ramps of the next step read) and goes on. wf_params::tuning bit 11
        // keeps everything in stream order (A/B measurement).
        wf_ctx::Comm &m = c->comm; const int p = m.red_parity; m.red_parity ^= 1;
        if (m.red_pending[p]) { WF_CUDA(c, cudaStreamWaitEvent(c->stream, m.ev_red_done[p], 0)); m.red_pending[p] = false; }     // buffer set of two steps ago
        k_reduce_rank<<<1, 256, 0, c->stream>>>(c->d_partial, nb, m.d_red_rank[p], c->d_mf_part, c->d_mf_flag, c->mf_blocks, (int)std::min<size_t>(c->meshes.size(), WF_MAX_MESH_FORCES)); WF_LAUNCH_CHECK(c);
        k_advance_time<<<1, 1, 0, c->stream>>>(c->d_sc, c->dt); WF_LAUNCH_CHECK(c);
        WF_CUDA(c, cudaEventRecord(m.ev_red_ready, c->stream));
        WF_CUDA(c, cudaStreamWaitEvent(m.stream, m.ev_red_ready, 0));
        WF_CHECK(comm_gather_scalars_on(c, m.d_red_rank[p], m.d_red_all[p], 32, m.stream));
        k_reduce_final<<<1, 32, 0, m.stream>>>(m.d_red_all[p], m.nranks, c->d_sc, c->dt, 2); WF_LAUNCH_CHECK(c);
        WF_CUDA(c, cudaEventRecord(m.ev_red_done[p], m.stream));
        m.red_pending[p] = true;
        c->launches += 4;
        return 0;
    }
    WF_CHECK(comm_red_wait(c));
    k_reduce_rank<<<1, 256, 0, c->stream>>>(c->d_partial, nb, c->comm.d_rank, c->d_mf_part, c->d_mf_flag, c->mf_blocks, (int)std::min<size_t>(c->meshes.size(), WF_MAX_MESH_FORCES)); WF_LAUNCH_CHECK(c);
    WF_CHECK(comm_gather_scalars(c, 32));
    k_reduce_final<<<1, 32, 0, c->stream>>>(multi ? c->comm.d_all : c->comm.d_rank, multi ? c->comm.nranks : 1, c->d_sc, c->dt, accumulate ? 1 : 0); WF_LAUNCH_CHECK(c);
    c->launches += 3;
    return 0;
}

int step_end_impl(wf_ctx *c)
{
    // end-of-step reductions + Time += dt on the device
    WF_CHECK(reduce_impl(c, true));
    c->time += c->dt;
    if (c->prm.contact && !c->meshes.empty()) {
        ProfScope ps(c, PK_END);
        // meshes are concatenated in the device arrays: one small launch per mesh for the nodes, one for all elements
        for (size_t m = 0; m < c->meshes.size(); ++m) {
            const HostMesh &hm = c->meshes[m];
            k_mesh_nodes_one<<<grid_for(hm.n_nodes, 128), 128, 0, c->stream>>>(hm.n_nodes, c->dm.node + 3 * hm.node0, c->dm.node_v + 3 * hm.node0, c->dm.motion + 16 * m, c->dt);
            WF_LAUNCH_CHECK(c); c->launches++;
        }
        k_mesh_elems<<<grid_for(c->dm.n_elems, 128), 128, 0, c->stream>>>(c->dm.n_elems, (int)(c->n_loc - c->nmeshp), c->dm.node, c->dm.node_v, c->dm.en, c->dm.nfar, c->dm.enormal, c->dm.ev, c->dm.pplane,
            c->slot_of, c->Q1[c->cur], c->F[c->cur], c->cap);
        WF_LAUNCH_CHECK(c); c->launches++;
    }
    return 0;
}

} // namespace wf

// ================================================================================================================
// C ABI: phases and whole steps
// ================================================================================================================
#define WF_NEED_READY(c) do { if (!(c)) return WF_ERR_ARG; if (!(c)->solver_ready) return wf::fail((c), WF_ERR_ARG, "%s: call wf_solver_init first", __func__); cudaSetDevice((c)->device); } while (0)

extern "C" int wf_solver_init(wf_ctx *c, double dt)
{
    if (!c) return WF_ERR_ARG;
    if (!c->have_particles) return fail(c, WF_ERR_ARG, "wf_solver_init: no particles");
    if (!(dt > 0.0)) return fail(c, WF_ERR_ARG, "wf_solver_init: dt must be > 0");
    cudaSetDevice(c->device);
    if (c->n_ghost > 0 && c->prm.solver == 3) return fail(c, WF_ERR_UNSUPPORTED, "ghost (symmetry) particles in ThermalSolve: the reference's loop never calls MoveGhost; not implemented");
    if (c->comm.nranks > 1) {         // "uniform over the particles" must mean over the DOMAIN: halo neighbours come from other ranks
        double mn = c->mass_min, mx = c->mass_max; WF_CHECK(comm_global_minmax(c, &mn, &mx)); c->uniform_mass = (mn == mx); c->mass_min = mn; c->mass_max = mx;
        double hn = c->uniform_h ? c->h_uniform : 0.0, hx = c->uniform_h ? c->h_uniform : 1.0e300; WF_CHECK(comm_global_minmax(c, &hn, &hx)); c->uniform_h = (hn == hx);
    }
    c->dt_init = c->dt_min = dt;
    c->dt = dt; c->time = 0.0; c->ts_i = 0; c->isfirst = true; c->isyielding = false; c->steps = 0; c->dens_valid = false; c->in_step = false;
    WF_CUDA(c, cudaMemsetAsync(c->d_sc, 0, sizeof(DevScalars), c->stream));      // a second wf_solver_init starts from zeroed energy / work accumulators like the first
    k_set_time<<<1, 1, 0, c->stream>>>(c->d_sc, 0.0, dt); WF_LAUNCH_CHECK(c);
    StepArgs a = make_step_args(c);
    k_init_pressure<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c);       // InitialChecks
    WF_CHECK(nb_cell_initiate(c));                                                               // CellInitiate (+ListGenerate: part of the build)
    if (c->prm.solver == 3) {
        // Domain::ThermalSolve (Thermal.cpp:567-626): one neighbour search for the whole run, the correction matrices, the initial dT/dt
        if (c->prm.contact) return fail(c, WF_ERR_UNSUPPORTED, "ThermalSolve with rigid-mesh contact is not implemented");
        WF_CHECK(nb_build_impl(c));
        c->solver_ready = true;
        WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1 | HALO_Q2));      // several GPUs: x, m/rho, T of the halo particles for the two walks below
        WF_CHECK(pair_gradcorr(c));
        WF_CHECK(pair_temp_inc_pp(c));
        WF_CHECK(reduce_impl(c, false));
        return 0;
    }
    if (c->prm.contact) {
        // SolverFraser.cpp:71-92: ID_orig = ID, MainNeighbourSearch, SaveNeighbourData, CalculateSurface(1)
        WF_CHECK(nb_build_impl(c));
    }
    c->nb_cleared = true;                                                                        // ClearNbData (:94)
    c->solver_ready = true;
    WF_CHECK(reduce_impl(c, false));
    return 0;
}

extern "C" int wf_nb_build(wf_ctx *c) { WF_NEED_READY(c); return nb_build_impl(c); }
extern "C" int wf_grad_corr(wf_ctx *c) { WF_NEED_READY(c); return pair_gradcorr(c); }
extern "C" int wf_density(wf_ctx *c) { WF_NEED_READY(c); if (!c->dens_valid) WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1)); return pair_density(c); }
extern "C" int wf_rate_stress(wf_ctx *c) { WF_NEED_READY(c); WF_CHECK(comm_halo_refresh(c, HALO_Q1)); return pair_rate_stress(c); }
extern "C" int wf_accel(wf_ctx *c) { WF_NEED_READY(c); WF_CHECK(comm_halo_refresh(c, HALO_Q2 | HALO_S)); return pair_accel(c); }
extern "C" int wf_contact(wf_ctx *c) { WF_NEED_READY(c); return contact_impl(c); }
extern "C" int wf_apply_bcs(wf_ctx *c) { WF_NEED_READY(c); return apply_bcs_impl(c); }
extern "C" int wf_integrate_pos(wf_ctx *c) { WF_NEED_READY(c); return move_pos_impl(c); }
extern "C" int wf_integrate_vel(wf_ctx *c) { WF_NEED_READY(c); return move_vel_impl(c); }
extern "C" int wf_thermal(wf_ctx *c) { WF_NEED_READY(c); return pair_thermal(c); }
extern "C" int wf_energy(wf_ctx *c) { WF_NEED_READY(c); return pair_energy(c); }

static int fetch_scalars(wf_ctx *c)
{
    WF_CHECK(comm_red_wait(c));
    WF_CUDA(c, cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int wf_reduce(wf_ctx *c, wf_scalars *out)
{
    WF_NEED_READY(c);
    if (!out) return WF_ERR_ARG;
    // the device scalars are refreshed by every step end (and by wf_solver_init); this is a read-back only
    WF_CHECK(fetch_scalars(c));
    const DevScalars &d = *c->h_sc;
    std::memset(out, 0, sizeof(*out));
    out->time = d.time; out->max_pl_strain = d.max_pl_strain;
    // Domain::max_disp as the reference holds it at this point: computed at the start of the step in flight / last completed
    for (int k = 0; k < 3; ++k) out->max_disp[k] = std::sqrt((c->in_step || c->steps == 0) ? d.max_disp2[k] : d.max_disp2_prev[k]);      // SolverFraser.cpp:159
    // the KickDrift / LeapFrog loops keep max_disp in a local that shadows Domain::max_disp (SolverKickDrift.cpp:122, SolverLeapfrog.cpp:138):
    // the member a driver reads is never written by them
    if (c->prm.solver != 0) for (int k = 0; k < 3; ++k) out->max_disp[k] = 0.0;
    out->kin_energy_sum = d.kin_energy_sum; out->int_energy_sum = d.int_energy_sum; out->plastic_work = d.plastic_work;
    out->contact_force_sum = d.contact_force_sum; out->contact_reaction_sum = d.contact_reaction_sum; out->ext_forces_work = d.ext_forces_work;
    out->contact_friction_work = d.contact_friction_work; out->accum_cont_heat_cond = d.accum_cont_heat_cond;
    out->max_T = d.max_T; out->min_T = d.min_T; out->steps = c->steps; out->nb_rebuilds = c->nb_rebuilds; out->is_yielding = c->isyielding ? 1 : 0;
    out->deltat = c->dt;
    return 0;
}

extern "C" int wf_mesh_forces(wf_ctx *c, double *out, int cap_meshes)
{
    WF_NEED_READY(c);
    if (!out || cap_meshes < 0) return WF_ERR_ARG;
    if ((int)c->meshes.size() > WF_MAX_MESH_FORCES) return fail(c, WF_ERR_UNSUPPORTED, "wf_mesh_forces: more than %d rigid meshes", WF_MAX_MESH_FORCES);
    WF_CHECK(fetch_scalars(c));
    for (int m = 0; m < cap_meshes; ++m) for (int d = 0; d < 3; ++d) out[3 * m + d] = (m < (int)c->meshes.size()) ? c->h_sc->mcf[m][d] : 0.0;
    return 0;
}

// SolverFraser.cpp:139-219: max pl_strain / max_disp and the neighbour-rebuild policy
static int nb_policy(wf_ctx *c)
{
    const double MIN_PS_FOR_NBSEARCH = 1.e-6;     // Domain.cpp:28
    // max pl_strain / max_disp as of the end of the previous step (device scalars are refreshed by every step end).
    // Once the body is yielding the cadence is fixed (pl_strain never decreases), so no host round trip is needed.
    double mx = MIN_PS_FOR_NBSEARCH * 2.0; bool check_every = false;
    if (!c->isyielding) {
        WF_CHECK(fetch_scalars(c));
        mx = c->h_sc->max_pl_strain;
        const V3 md = v3(std::sqrt(c->h_sc->max_disp2[0]), std::sqrt(c->h_sc->max_disp2[1]), std::sqrt(c->h_sc->max_disp2[2]));
        check_every = norm(md) > 0.1 * c->hmax;
    }
    c->step_max_pl = mx; c->in_step = true;
    if (mx > MIN_PS_FOR_NBSEARCH && !c->isyielding) {          // first time yielding (:172-179)
        WF_CHECK(nb_build_impl(c));
        c->isyielding = true;
    }
    if (mx > MIN_PS_FOR_NBSEARCH || c->isfirst || check_every) {
        if (c->ts_i == 0 && c->nb_cleared) {
            WF_CHECK(nb_build_impl(c));
            // SolverFraser.cpp:203-204, SolverLeapfrog.cpp:196-197: at the cadence rebuilds only, not at the first-yield search above;
            // the KickDrift loop computes the matrices once, at the first step (SolverKickDrift.cpp:187-190)
            if (c->prm.solver != 1 || c->isfirst) WF_CHECK(pair_gradcorr(c));
        }
    }
    return 0;
}
// SolverFraser.cpp:384-412: Time += dt, mesh update, neighbour cadence bookkeeping
static int finish_step(wf_ctx *c)
{
    const double MIN_PS_FOR_NBSEARCH = 1.e-6;
    WF_CHECK(step_end_impl(c));
    if (c->step_max_pl > MIN_PS_FOR_NBSEARCH) {                // :397-406
        if (c->ts_i == (c->prm.ts_nb_inc - 1)) c->nb_cleared = true;      // ClearNbData
        c->ts_i++;
        if (c->ts_i > (c->prm.ts_nb_inc - 1)) c->ts_i = 0;
    }
    c->isfirst = false; c->in_step = false;
    c->steps++;
    return 0;
}
extern "C" int wf_step_begin(wf_ctx *c) { WF_NEED_READY(c); return nb_policy(c); }   // StartAcceleration's TIR is fused into wf_density; its resets are implicit
extern "C" int wf_step_end(wf_ctx *c) { WF_NEED_READY(c); return finish_step(c); }

extern "C" int wf_fraser_stage_a(wf_ctx *c)
{
    WF_NEED_READY(c);
    WF_CHECK(nb_policy(c));
    if (!c->dens_valid) WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1));   // stand-alone density pass gathers x, v, m/rho of the halos
    WF_CHECK(pair_density(c));
    // each halo refresh travels on the communication stream while the interior slots are evaluated; the boundary layers follow
    int64_t i0 = 0, i1 = c->n_loc;
    WF_CHECK(comm_halo_begin(c, HALO_Q1, &i0, &i1));             // m/rho of the halos (rho just advanced)
    WF_CHECK(pair_rate_stress(c, i0, i1));
    WF_CHECK(comm_halo_end(c));
    WF_CHECK(pair_rate_stress(c, 0, i0)); WF_CHECK(pair_rate_stress(c, i1, c->n_loc));
    WF_CHECK(comm_halo_begin(c, HALO_Q2 | HALO_S, &i0, &i1));    // m, rho, sigma/rho^2, TIR of the halos
    WF_CHECK(comm_time_begin(c));                                // this rank's speed, for the slab rebalancing at the next rebuild
    WF_CHECK(pair_accel(c, i0, i1));
    WF_CHECK(comm_time_end(c));
    WF_CHECK(comm_halo_end(c));
    WF_CHECK(pair_accel(c, 0, i0)); WF_CHECK(pair_accel(c, i1, c->n_loc));
    WF_CHECK(contact_impl(c));
    return 0;
}
extern "C" int wf_fraser_stage_b1(wf_ctx *c) { WF_NEED_READY(c); WF_CHECK(move_ghost_impl(c)); return move_pos_impl(c); }      // MoveGhost (:335), x update
extern "C" int wf_fraser_stage_b2(wf_ctx *c) { WF_NEED_READY(c); WF_CHECK(move_vel_impl(c)); return move_ghost_impl(c); }      // v update, MoveGhost (:358)
extern "C" int wf_move_ghost(wf_ctx *c) { WF_NEED_READY(c); return move_ghost_impl(c); }
extern "C" int wf_fraser_stage_c(wf_ctx *c)
{
    WF_NEED_READY(c);
    // CalcPlasticWorkHeat + ThermalCalcs + CalcKinEnergyEqn + CalcIntEnergyEqn (+ next density rate, unless the cadence
    // (finish_step / nb_policy) says the next step starts with a neighbour rebuild, which would invalidate it)
    const bool next_rebuilds = c->step_max_pl > 1.e-6 && c->ts_i == (c->prm.ts_nb_inc - 1);
    WF_CHECK(pair_post(c, next_rebuilds));
    return finish_step(c);
}

// Domain::AdaptiveTimeStep (Domain.cpp:205-239), host scalars; auto_ts_cont is off in every driver and not modelled
static void adaptive_time_step(wf_ctx *c)
{
    const double manual_min_ts = 1.e-10;       // Domain.cpp:170
    double &deltat = c->dt; const double deltatint = c->dt_init, deltatmin = c->dt_min;
    if (deltatint > deltatmin) {
        if (deltat < deltatmin) deltat = 2.0 * deltat * deltatmin / (deltat + deltatmin);
        else deltat = deltatmin;
    } else {
        if (deltatint != deltat) deltat = 2.0 * deltat * deltatint / (deltat + deltatint);
        else deltat = deltatint;
    }
    if (deltat < manual_min_ts) deltat = manual_min_ts;
}

static int kick(wf_ctx *c, int kind, double prev_dt)
{
    ProfScope ps(c, PK_MOVE);
    StepArgs a = make_step_args(c);
    k_kick<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a, kind, prev_dt, c->dt); WF_LAUNCH_CHECK(c); c->launches++;
    c->dens_valid = false;
    return 0;
}

// One step of SolveDiffUpdateKickDrift (SolverKickDrift.cpp:98-335) or SolveDiffUpdateLeapFrog (SolverLeapfrog.cpp:124-443), in five
// stages split at the loops' GeneralAfter calls: same kernels as the Fraser step in the order those loops run them -- momentum
// first (with the TIR of the state the previous step left), density after the first kick, rate tensors and stresses after the drift.
//   0  rebuild policy, StartAcceleration (TIR), CalcAccel                       [GeneralAfter]
//   1  contact; LeapFrog, first step: v -= a*dt/2                               [GeneralAfter, LeapFrog only]
//   2  CheckMinTSVel + AdaptiveTimeStep, velocity kick                          [GeneralAfter]
//   3  density, drift; KickDrift: second half kick                              [GeneralAfter, KickDrift only]
//   4  rate tensors + stresses, heat, energies, end of step
// device_bcs: apply the device-side GeneralAfter at those points (wf_run); the host mirror calls the driver's callback instead.
static int kdlf_stage(wf_ctx *c, int stage, bool device_bcs)
{
    const int solver = c->prm.solver;
    switch (stage) {
    case 0:
        WF_CHECK(nb_policy(c));
        WF_CHECK(pair_tir(c));
        WF_CHECK(comm_halo_refresh(c, HALO_Q2 | HALO_S));
        WF_CHECK(pair_accel(c));
        if (device_bcs) WF_CHECK(apply_bcs_impl(c));
        return 0;
    case 1:
        WF_CHECK(contact_impl(c));
        if (solver == 2) {
            if (c->isfirst) WF_CHECK(kick(c, 0, c->dt));
            WF_CHECK(move_ghost_impl(c));                            // SolverLeapfrog.cpp:253
            if (device_bcs) WF_CHECK(apply_bcs_impl(c));
        }
        return 0;
    case 2:
        c->prev_dt = c->dt;
        if (c->prm.auto_ts || c->prm.auto_ts_acc) {
            double t = 0.0;
            WF_CHECK(pair_min_ts(c, &t));
            if (c->comm.nranks > 1) {                                // min over the ranks
                WF_CUDA(c, cudaMemcpyAsync(c->comm.d_rank, &t, sizeof(double), cudaMemcpyHostToDevice, c->stream));
                WF_CHECK(comm_gather_scalars(c, 1));
                std::vector<double> all(c->comm.nranks);
                WF_CUDA(c, cudaMemcpyAsync(all.data(), c->comm.d_all, all.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                WF_CUDA(c, cudaStreamSynchronize(c->stream));
                for (double v : all) if (v < t) t = v;
            }
            c->dt_min = c->dt_init;                                  // CheckMinTSVel: deltatmin = deltatint; if (deltatmin > test) deltatmin = test
            if (c->dt_min > t) c->dt_min = t;
            adaptive_time_step(c);
        }
        WF_CHECK(kick(c, solver == 1 ? 1 : 2, c->prev_dt));
        WF_CHECK(move_ghost_impl(c));                                // SolverKickDrift.cpp:230, SolverLeapfrog.cpp:272
        if (device_bcs) WF_CHECK(apply_bcs_impl(c));
        return 0;
    case 3:
        WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1));
        c->dens_valid = false;
        WF_CHECK(pair_density(c, false));                            // CalcDensInc + DensReduction, rho += dt*drho; the step's TIR stands (pair_tir, stage 0)
        WF_CHECK(kick(c, 4, c->prev_dt));                            // drift
        if (solver == 1) { WF_CHECK(kick(c, 3, c->prev_dt)); WF_CHECK(move_ghost_impl(c)); if (device_bcs) WF_CHECK(apply_bcs_impl(c)); }   // SolverKickDrift.cpp:266
        return 0;
    case 4:
        WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1));
        WF_CHECK(pair_rate_stress(c));
        WF_CHECK(pair_post(c, true));                                // CalcPlasticWorkHeat, ThermalCalcs, energies (no density hand-over in this step order)
        c->dens_valid = false;
        return finish_step(c);
    default:
        return fail(c, WF_ERR_ARG, "wf_solver_stage: stage %d (0..4)", stage);
    }
}
static int kdlf_step(wf_ctx *c)
{
    for (int st = 0; st < 5; ++st) WF_CHECK(kdlf_stage(c, st, true));
    return 0;
}
extern "C" int wf_solver_stage(wf_ctx *c, int stage)
{
    WF_NEED_READY(c);
    if (c->prm.solver == 0) return fail(c, WF_ERR_ARG, "wf_solver_stage: the Fraser step has its own stages (wf_fraser_stage_a/b1/b2/c)");
    if (c->prm.solver == 3) return fail(c, WF_ERR_ARG, "wf_solver_stage: ThermalSolve steps are wf_thermal + the callback + wf_step_end; use wf_run");
    return kdlf_stage(c, stage, false);
}

// One step of Domain::ThermalSolve (Thermal.cpp:627-717): ThermalCalcs (T += dt*dTdt, CalcConvHeat, CalcTempInc, thermal-expansion
// strain rate), GeneralAfter, Time += dt; the neighbour lists of the initial search are kept
static int thermal_step(wf_ctx *c)
{
    WF_CHECK(pair_thermal(c));
    WF_CHECK(apply_bcs_impl(c));
    WF_CHECK(reduce_impl(c, true));
    c->time += c->dt; c->steps++;
    return 0;
}
static int one_step(wf_ctx *c)
{
    if (c->prm.solver == 3) return thermal_step(c);
    if (c->prm.solver != 0) return kdlf_step(c);
    return wf_fraser_run(c, 1);
}
extern "C" int wf_run(wf_ctx *c, int64_t nsteps)
{
    WF_NEED_READY(c);
    if (c->prm.solver == 0) return wf_fraser_run(c, nsteps);
    for (int64_t k = 0; k < nsteps; ++k) WF_CHECK(one_step(c));
    return 0;
}
extern "C" int wf_run_until(wf_ctx *c, double tf, int64_t *steps_done)
{
    WF_NEED_READY(c);
    int64_t n = 0;
    // while (Time <= tf) in the Fraser / KickDrift / LeapFrog loops (SolverFraser.cpp:133); ThermalSolve stops at Time < tf (Thermal.cpp:626)
    while (c->prm.solver == 3 ? c->time < tf : c->time <= tf) { WF_CHECK(one_step(c)); ++n; }
    if (steps_done) *steps_done = n;
    return 0;
}

// nsteps whole steps with the device-side boundary conditions at both GeneralAfter points; no host callback.
extern "C" int wf_fraser_run(wf_ctx *c, int64_t nsteps)
{
    WF_NEED_READY(c);
    for (int64_t k = 0; k < nsteps; ++k) {
        WF_CHECK(wf_fraser_stage_a(c));
        // GeneralAfter, x update, v update, GeneralAfter fused in one pass (per-particle, order preserved)
        {
            ProfScope ps(c, PK_MOVE);
            StepArgs a = make_step_args(c);
            if (!c->meshes.empty()) { k_mesh_ramp<<<1, 32, 0, c->stream>>>(c->dm.motion, (int)c->meshes.size(), c->d_sc); WF_LAUNCH_CHECK(c); c->launches++; }
            if (c->n_ghost == 0) { k_bc_move<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a, 15); WF_LAUNCH_CHECK(c); c->launches++; }
            else {
                // with ghost pairs the pass splits at the two MoveGhost calls (SolverFraser.cpp:335,358)
                const int g = grid_for(c->n_loc, 256);
                k_bc_move<<<g, 256, 0, c->stream>>>(a, 1); WF_LAUNCH_CHECK(c); c->launches++;
                WF_CHECK(move_ghost_impl(c));
                k_bc_move<<<g, 256, 0, c->stream>>>(a, 2 | 4); WF_LAUNCH_CHECK(c); c->launches++;
                WF_CHECK(move_ghost_impl(c));
                k_bc_move<<<g, 256, 0, c->stream>>>(a, 8); WF_LAUNCH_CHECK(c); c->launches++;
            }
        }
        WF_CHECK(wf_fraser_stage_c(c));
    }
    return 0;
}
