// bgs_pybind.cpp -- a C++ Python extension over the C ABI (include/bgs_b200.h), built with pybind11.
//
// This is the binding a maintainer of the reference would put next to
// src/simulator/game/connect.cpp:19-62 / bounce.cpp:19-61 (there with nanobind, which is not installed
// here; pybind11 is): it proves the drop-in boundary from C++ -- plain pointers and sizes, no torch
// types -- and is exercised by tests/test_gpu_pybind.py.  The shipped host API (simulator.game.*,
// simulator.batch) calls the same entry points through ctypes.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/bgs_b200.h"

namespace py = pybind11;

template <class T>
using Arr = py::array_t<T, py::array::c_style>;

// -> Python RuntimeError, the exception textual/connect.py:115-118 and textual/bounce.py:118-128 expect
static void check(int rc) {
    if (rc != BGS_OK) throw std::runtime_error(std::string(bgs_last_error()) + " (code " + std::to_string(rc) + ")");
}

template <class T>
static T* data_or_null(py::object o, py::ssize_t need, const char* name) {
    if (o.is_none()) return nullptr;
    auto a = o.cast<Arr<T>>();
    if (a.size() != need) throw py::type_error(std::string(name) + ": wrong size");  // the caster's shape check, tensor.hpp:45-59
    if (!a.writeable()) throw py::type_error(std::string(name) + ": array is read-only");
    return a.mutable_data();
}

template <class T>
static T* dev(std::uintptr_t p) { return reinterpret_cast<T*>(p); }

PYBIND11_MODULE(_bgs_pybind, m) {
    m.doc() = "pybind11 binding of libbgs_b200.so (include/bgs_b200.h)";
    m.attr("STATS_LEN") = BGS_STATS_LEN;
    m.def("version", &bgs_version);
    m.def("device_count", &bgs_device_count);
    m.def("connect_supported", &bgs_connect_supported);

    // Config(h, w, k).rollout(...): the README.md:49-72 loop for n games, host buffers in and out
    m.def("connect_rollout",
          [](int h, int w, int k, std::uint64_t n, std::uint64_t game_id0, std::uint64_t seed, py::object actions,
             py::object length, py::object winner, py::object final_grid, py::object reward, py::object stats, int device) {
              const py::ssize_t hw = (py::ssize_t)h * w, ns = (py::ssize_t)n;
              std::uint8_t* a = data_or_null<std::uint8_t>(actions, ns * hw, "actions");
              std::uint8_t* l = data_or_null<std::uint8_t>(length, ns, "length");
              std::int8_t* wi = data_or_null<std::int8_t>(winner, ns, "winner");
              std::int8_t* g = data_or_null<std::int8_t>(final_grid, ns * hw, "final_grid");
              float* r = data_or_null<float>(reward, ns * 2, "reward");
              std::int64_t* s = data_or_null<std::int64_t>(stats, BGS_STATS_LEN, "stats");
              py::gil_scoped_release release;  // the reference never releases the GIL; this call can
              check(bgs_connect_rollout_host(device, h, w, k, n, game_id0, seed, a, l, wi, g, r, s));
          },
          py::arg("height"), py::arg("width"), py::arg("count"), py::arg("n_games"), py::arg("game_id0"), py::arg("seed"),
          py::arg("actions") = py::none(), py::arg("length") = py::none(), py::arg("winner") = py::none(),
          py::arg("final_grid") = py::none(), py::arg("reward") = py::none(), py::arg("stats") = py::none(),
          py::arg("device") = 0);

    m.def("bounce_rollout",
          [](Arr<std::int8_t> grid0, int rules, int max_plies, std::uint64_t n, std::uint64_t game_id0, std::uint64_t seed,
             py::object moves, py::object length, py::object winner, py::object final_grid, py::object reward,
             py::object stats, int device) {
              if (grid0.ndim() != 2) throw py::type_error("grid0 must be a 2-D int8 array");
              const int h = (int)grid0.shape(0), w = (int)grid0.shape(1);
              const py::ssize_t hw = (py::ssize_t)h * w, ns = (py::ssize_t)n;
              std::uint8_t* mv = data_or_null<std::uint8_t>(moves, ns * max_plies * 2, "moves");
              std::uint16_t* l = data_or_null<std::uint16_t>(length, ns, "length");
              std::int8_t* wi = data_or_null<std::int8_t>(winner, ns, "winner");
              std::int8_t* g = data_or_null<std::int8_t>(final_grid, ns * hw, "final_grid");
              float* r = data_or_null<float>(reward, ns * 2, "reward");
              std::int64_t* s = data_or_null<std::int64_t>(stats, BGS_STATS_LEN, "stats");
              const std::int8_t* g0 = grid0.data();
              py::gil_scoped_release release;
              check(bgs_bounce_rollout_host(device, g0, h, w, rules, max_plies, n, game_id0, seed, mv, l, wi, g, r, s));
          },
          py::arg("grid0"), py::arg("rules"), py::arg("max_plies"), py::arg("n_games"), py::arg("game_id0"), py::arg("seed"),
          py::arg("moves") = py::none(), py::arg("length") = py::none(), py::arg("winner") = py::none(),
          py::arg("final_grid") = py::none(), py::arg("reward") = py::none(), py::arg("stats") = py::none(),
          py::arg("device") = 0);

    // device-pointer variants for callers that already hold CUDA memory (DLPack / torch / cupy): raw addresses
    m.def("connect_rollout_device",
          [](int h, int w, int k, std::uint64_t n, std::uint64_t id0, std::uint64_t seed, std::uintptr_t actions,
             std::uintptr_t length, std::uintptr_t winner, std::uintptr_t final_grid, std::uintptr_t reward,
             std::uintptr_t stats, std::uintptr_t stream) {
              check(bgs_connect_rollout_export(h, w, k, n, id0, seed, dev<std::uint8_t>(actions), dev<std::uint8_t>(length),
                                               dev<std::int8_t>(winner), dev<std::int8_t>(final_grid), dev<float>(reward),
                                               dev<std::int64_t>(stats), dev<void>(stream)));
          });
    // State::get_action_at + Action::sample_next_state + has_ended / reward / actions of the new state, batched
    m.def("connect_step_device",
          [](int h, int w, int k, std::uint64_t n, std::uintptr_t grid, std::uintptr_t player, std::uintptr_t winner,
             std::uintptr_t action, std::uintptr_t grid_out, std::uintptr_t player_out, std::uintptr_t winner_out,
             std::uintptr_t ended_out, std::uintptr_t reward_out, std::uintptr_t legal_out, std::uintptr_t status,
             std::uintptr_t stream) {
              check(bgs_connect_step(h, w, k, n, dev<const std::int8_t>(grid), dev<const std::int8_t>(player),
                                     dev<const std::int8_t>(winner), dev<const std::int32_t>(action), dev<std::int8_t>(grid_out),
                                     dev<std::int8_t>(player_out), dev<std::int8_t>(winner_out), dev<std::uint8_t>(ended_out),
                                     dev<float>(reward_out), dev<std::uint32_t>(legal_out), dev<std::int32_t>(status),
                                     dev<void>(stream)));
          });
    m.def("bounce_step_device",
          [](int h, int w, int rules, std::uint64_t n, std::uintptr_t grid, std::uintptr_t player, std::uintptr_t winner,
             std::uintptr_t ended, std::uintptr_t move, std::uintptr_t grid_out, std::uintptr_t player_out,
             std::uintptr_t winner_out, std::uintptr_t ended_out, std::uintptr_t reward_out, std::uintptr_t status,
             std::uintptr_t stream) {
              check(bgs_bounce_step(h, w, rules, n, dev<const std::int8_t>(grid), dev<const std::int8_t>(player),
                                    dev<const std::int8_t>(winner), dev<const std::uint8_t>(ended), dev<const std::int32_t>(move),
                                    dev<std::int8_t>(grid_out), dev<std::int8_t>(player_out), dev<std::int8_t>(winner_out),
                                    dev<std::uint8_t>(ended_out), dev<float>(reward_out), dev<std::int32_t>(status),
                                    dev<void>(stream)));
          });
}
