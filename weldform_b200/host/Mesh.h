// Host-side mirror of SPH::TriMesh / SPH::Element (Source/Mesh.h:44-75, Mesh.cpp): rigid tool surfaces. The driver owns
// the TriMesh; Domain::AddTrimeshParticles hands its nodes / elements / normals / radii to the device (wf_add_mesh),
// where TriMesh::Update runs every step (wf_step.cu k_mesh_nodes_one / k_mesh_elems). SetVel / SetRotAxisVel called from
// a GeneralAfter callback are forwarded with wf_mesh_set_motion. After Solve the node positions are read back.
#ifndef WF_HOST_MESH_H
#define WF_HOST_MESH_H
#include "matvec.h"
#include "NastranReader.h"

namespace SPH {

class Element {
public:
    Element() {}
    Element(const int & n1, const int & n2, const int & n3) { node[0] = n1; node[1] = n2; node[2] = n3; }
    Vec3_t centroid, normal, v;
    double radius = 0.0, pplane = 0.0, area = 0.0;
    int node[3] = { 0, 0, 0 }, nfar = 0;
};

class TriMesh {
public:
    Array<Element *> element;
    Array<Vec3_t *> node, node_v;
    int id = 0, dimension = 3;
    Vec3_t m_v, m_w;
    double T = 0.0;
    TriMesh() {}
    // Source/Mesh.cpp:17-60
    TriMesh(NastranReader & nr, bool flipnormals = false)
    {
        dimension = nr.dim;
        for (int n = 0; n < nr.node_count; ++n) { node.Push(new Vec3_t(nr.node[3 * n], nr.node[3 * n + 1], nr.node[3 * n + 2])); node_v.Push(new Vec3_t(0., 0., 0.)); }
        for (int e = 0; e < nr.elem_count; ++e) {
            const int n0 = nr.elcon[3 * e], n1 = nr.elcon[3 * e + 1], n2 = nr.elcon[3 * e + 2];
            element.Push(flipnormals ? new Element(n1, n0, n2) : new Element(n0, n1, n2));
            element[e]->centroid = (*node[n0] + *node[n1] + *node[n2]) / 3.;
            Vec3_t w = cross(*node[n1] - *node[n0], *node[n2] - *node[n0]);
            element[e]->normal = w / norm(w);
        }
    }
    ~TriMesh() { for (size_t i = 0; i < element.Size(); ++i) delete element[i]; for (size_t i = 0; i < node.Size(); ++i) { delete node[i]; delete node_v[i]; } }
    // Source/Mesh.cpp:82-202 (3-D): (dens+1)^2 nodes on the plane {x_axis = p1[axis]}, 2*dens^2 triangles, normals flipped for positaxisorent
    inline void AxisPlaneMesh(const int & axis, bool positaxisorent, const Vec3_t p1, const Vec3_t p2, const int & dens)
    {
        int dir[2];
        if (axis == 0) { dir[0] = 1; dir[1] = 2; } else if (axis == 1) { dir[0] = 0; dir[1] = 2; } else { dir[0] = 0; dir[1] = 1; }
        const double x3 = p1(axis); double x2 = p1(dir[1]);
        const double dl = (p2(dir[0]) - p1(dir[0])) / dens;
        for (int j = 0; j < dens + 1; ++j) {
            double x1 = p1(dir[0]);
            for (int i = 0; i < dens + 1; ++i) {
                Vec3_t * v = new Vec3_t(); (*v)(dir[0]) = x1; (*v)(dir[1]) = x2; (*v)(axis) = x3;
                node.Push(v); node_v.Push(new Vec3_t(0., 0., 0.));
                x1 += dl;
            }
            x2 += dl;
        }
        for (int j = 0; j < dens; ++j) for (int i = 0; i < dens; ++i) {
            const int n0 = (dens + 1) * j + i, n1 = n0 + 1, n2 = (dens + 1) * (j + 1) + i, n3 = n2 + 1;
            if (positaxisorent) { add_elem(n0, n1, n2); add_elem(n1, n3, n2); }
            else                { add_elem(n0, n2, n1); add_elem(n1, n2, n3); }
        }
        for (size_t e = 0; e < element.Size(); ++e) {
            Element * el = element[e];
            Vec3_t w = cross(*node[el->node[1]] - *node[el->node[0]], *node[el->node[2]] - *node[el->node[0]]);
            el->normal = w / norm(w);
            if (positaxisorent) el->normal = -el->normal;
        }
    }
    // Source/Mesh.cpp:206-228: bounding radius about the centroid; nfar ends as the last node (reference quirk)
    inline void CalcSpheres()
    {
        for (size_t e = 0; e < element.Size(); ++e) {
            Element * el = element[e]; double maxr = 0.;
            for (int n = 0; n < 3; ++n) { const double r = norm(*node[el->node[n]] - el->centroid); if (r > maxr) maxr = r; el->nfar = n; }
            el->radius = maxr;
        }
        UpdatePlaneCoeff();
    }
    inline void UpdatePlaneCoeff() { for (size_t e = 0; e < element.Size(); ++e) element[e]->pplane = dot(*node[element[e]->node[element[e]->nfar]], element[e]->normal); }
    inline void SetVel(const Vec3_t & v) { m_v = v; }
    inline void SetRotAxisVel(const Vec3_t & omega) { m_w = omega; }
    inline void Move(const Vec3_t & d) { for (size_t n = 0; n < node.Size(); ++n) *node[n] += d; for (size_t e = 0; e < element.Size(); ++e) element[e]->centroid += d; UpdatePlaneCoeff(); }
private:
    void add_elem(int a, int b, int c)
    {
        Element * el = new Element(a, b, c);
        el->centroid = ((*node[a] + *node[b]) + *node[c]) / 3.;
        element.Push(el);
    }
};

} // namespace SPH
#endif
