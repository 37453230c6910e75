// Test-infrastructure only.  A second, tiny pybind11 module compiled TOGETHER WITH the
// unmodified reference sources (see build_ref.sh).  It drives the reference Simulation
// class through its own methods and exposes internals that redmax_py does not bind:
// the active contact-point index lists of every Force, the marker->body ids of every
// tactile sensor, and the assembled M/K/D/H matrices.  Nothing here restates reference
// arithmetic; it only reads reference state.  Used by tests/golden/make_golden.py to
// create the golden fixtures and by developers to localise a parity failure.
#include <sstream>
#include <iostream>
#include <fstream>
#include <complex>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <memory>
#include <functional>
#include <algorithm>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#define private public
#define protected public
#include "Simulation.h"
#include "Body/BodyPrimitiveShape.h"
#include "Robot.h"
#include "Body/Body.h"
#include "Joint/Joint.h"
#include "Force/Force.h"
#include "Force/ForceGroundContact.h"
#include "Force/ForceGeneralPrimitiveContact.h"
#include "Sensor/TactileSensor.h"
#include "CollisionDetection/CollisionDetection.h"
#undef private
#undef protected
#include <pybind11/pybind11.h>
#include <pybind11/eigen.h>
#include <pybind11/stl.h>

namespace py = pybind11;
using namespace redmax;

// Counters incremented by the instrumented compile of Simulation::newton that build_ref.sh links into THIS module only
// (one line each, inserted by sed after the func_with_derivatives / func calls of DH/Simulation.cpp:1171,1189).
int g_probe_newton_iters = 0, g_probe_ls_evals = 0;

// (Newton iterations, line-search evaluations) since the last call; resets the counters.
static py::tuple newton_counts(Simulation&) {
    py::tuple t = py::make_tuple(g_probe_newton_iters, g_probe_ls_evals);
    g_probe_newton_iters = 0;
    g_probe_ls_evals = 0;
    return t;
}

static int body_index(Robot* robot, Body* b) {
    if (b == nullptr) return -1;
    for (size_t i = 0; i < robot->_bodies.size(); ++i)
        if (robot->_bodies[i] == b) return (int)i;
    return -2;
}

// contact index sets at the simulation's CURRENT state (after forward()).
static py::dict contact_sets(Simulation& sim) {
    py::dict out;
    py::list ground, gp, marker_body;
    Robot* robot = sim._robot;
    for (auto force : robot->_forces) {
        if (auto g = dynamic_cast<ForceGroundContact*>(force)) {
            std::vector<Contact> contacts;
            collision_detection_ground_object(g->_E_g, g->_contact_body, contacts);
            std::vector<int> ids;
            for (auto& c : contacts) ids.push_back(c._id);
            ground.append(ids);
        } else if (auto p = dynamic_cast<ForceGeneralPrimitiveContact*>(force)) {
            std::vector<Contact> contacts;
            collision_detection_general_primitive(p->_contact_body, p->_primitive_body, contacts);
            std::vector<int> ids;
            for (auto& c : contacts) ids.push_back(c._id);
            gp.append(ids);
        }
    }
    for (auto sensor : robot->_tactile_sensors) {
        sensor->compute_tactile_values();
        std::vector<int> ids;
        for (auto b : sensor->_contact_body) ids.push_back(body_index(robot, b));
        marker_body.append(ids);
    }
    out["ground"] = ground;
    out["gp"] = gp;
    out["marker_body"] = marker_body;
    return out;
}

// M, fr, K, D (+ H for a given q0/qdot0) evaluated by the reference at (q, qdot, u).
static py::dict matrices(Simulation& sim, VectorX q, VectorX qdot, VectorX u) {
    sim.set_u(u);
    sim.set_state(q, qdot);
    sim.update_robot();
    int n = sim._ndof_r;
    MatrixX M = MatrixX::Zero(n, n), K = MatrixX::Zero(n, n), D = MatrixX::Zero(n, n);
    VectorX fr = VectorX::Zero(n);
    JacobianMatrixVector dM_dq(n, n, n);
    sim.computeMatrices(M, fr, dM_dq, K, D);
    py::dict out;
    out["M"] = M; out["fr"] = fr; out["K"] = K; out["D"] = D;
    py::list dM;
    for (int k = 0; k < n; ++k) dM.append(MatrixX(dM_dq(k)));
    out["dM_dq"] = dM;
    out["J"] = sim._J; out["Jdot"] = sim._Jdot; out["fm"] = sim._fm; out["Km"] = sim._Km; out["Dm"] = sim._Dm;
    out["dphi_dq"] = sim._dphi_dq;
    return out;
}

// per-step tape saved by the reference in backward mode
static py::dict tape(Simulation& sim, int k) {
    py::dict out;
    out["M"] = sim._backward_info._M[k];
    out["D"] = sim._backward_info._D[k];
    out["H"] = MatrixX(sim._backward_info._H_lu[k].reconstructedMatrix());
    out["dg_du"] = sim._backward_info._dg_du[k];
    out["dvar_dq"] = sim._backward_info._dvar_dq[k];
    out["dtactile_dq"] = sim._backward_info._dtactile_dq[k];
    out["dtactile_dqdot"] = sim._backward_info._dtactile_dqdot[k];
    return out;
}

static py::list body_frames(Simulation& sim) {
    py::list out;
    for (auto b : sim._robot->_bodies) {
        py::dict d;
        d["E_0i"] = Matrix4(b->_E_0i);
        d["phi"] = VectorX(b->_phi);
        d["E_ji"] = Matrix4(b->_E_ji);
        d["inertia"] = VectorX(b->_Inertia);
        out.append(d);
    }
    return out;
}

PYBIND11_MODULE(redmax_probe, m) {
    m.doc() = "read-only probes into the reference Simulation (test infrastructure)";
    // Own binding of the class under another python name so the probe is self-contained
    // (redmax_py's binding lives in a different shared object).
    py::class_<Simulation>(m, "ProbeSimulation", py::module_local())
        .def(py::init<std::string, bool>(), py::arg("xml_file_path"), py::arg("verbose") = false)
        .def_readonly("ndof_r", &Simulation::_ndof_r)
        .def_readonly("ndof_u", &Simulation::_ndof_u)
        .def_readonly("ndof_var", &Simulation::_ndof_var)
        .def_readonly("ndof_tactile", &Simulation::_ndof_tactile)
        .def("set_state_init", &Simulation::set_state_init)
        .def("set_q_init", &Simulation::set_q_init)
        .def("get_q_init", &Simulation::get_q_init)
        .def("reset", &Simulation::reset, py::arg("backward_flag") = false, py::arg("backward_design_params_flag") = false)
        .def("set_u", &Simulation::set_u)
        .def("forward", &Simulation::forward, py::arg("num_steps"), py::arg("verbose") = false,
             py::arg("test_derivatives") = false, py::arg("save_last_frame_var_only") = false)
        .def("get_q", &Simulation::get_q)
        .def("get_qdot", &Simulation::get_qdot)
        .def("get_variables", &Simulation::get_variables)
        .def("get_tactile_force_vector", &Simulation::get_tactile_force_vector)
        .def("set_qdot_init", &Simulation::set_qdot_init)
        .def("newton_counts", &newton_counts)
        .def("contact_sets", &contact_sets)
        .def("matrices", &matrices)
        .def("tape", &tape)
        .def("body_frames", &body_frames);
}
