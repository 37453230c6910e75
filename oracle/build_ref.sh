#!/usr/bin/env bash
# Headless build of the UNMODIFIED reference DiffRedMax (externals/DiffHand/core/projects/redmax)
# as the python module `redmax_py`, from the sources where they lie under /root/reference.
# Outputs go ONLY into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# The reference's own CMake build is not used (it requires OpenGL/X11 dev packages that this
# image lacks); the viewer is replaced by a 6-line stub so Simulation::replay() is a no-op.
# Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
# --impl reference legs may load what this script builds.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF_ROOT:-/root/reference}"
C="$REF/externals/DiffHand/core"
OUT="$HERE/_ref"
PY="${PYTHON:-python3}"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
if [ ! -d "$C/projects/redmax" ]; then
  echo "build_ref: $C not present (GPU box?) - using prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/assets"
cat > "$OUT/defs.h" <<X
#define GRAPHICS_CODEBASE_SOURCE_DIR "$C"
#define PROJECT_DIR "$C"
X
cat > "$OUT/SimViewerStub.cpp" <<'X'
#include "SimViewer.h"
#include "Simulation.h"
namespace redmax {
SimViewer::SimViewer(Simulation* sim) : _sim(sim), _timer(nullptr), _keyboard_handler(nullptr) {}
SimViewer::~SimViewer() {}
void SimViewer::initialize() {}
void SimViewer::run() {}
}
X
SRCS="$(find "$C/projects/redmax" -name '*.cpp' ! -name main.cpp ! -name SimViewer.cpp ! -name Test.cpp | sort) \
 $C/projects/opengl_viewer/src/geometry.cpp $C/projects/opengl_viewer/src/image.cpp $C/projects/opengl_viewer/src/bounding_box.cpp \
 $C/externals/pugixml/src/pugixml.cpp $C/externals/tiny_obj_loader/tiny_obj_loader.cpp $OUT/SimViewerStub.cpp"
# Release flags of the reference's own build (core/setup.py:39,49 -> CMAKE_BUILD_TYPE=Release = -O3 -DNDEBUG)
FLAGS="-O3 -DNDEBUG -std=c++14 -fPIC -w -DGLEW_STATIC -DGLEW_NO_GLU -DGLFW_INCLUDE_NONE -include $OUT/defs.h \
 -I$C/projects/redmax -I$C/projects/opengl_viewer/include -I$C/externals/eigen -I$C/externals/glew/include \
 -I$C/externals/glfw/include/GLFW -I$C/externals/imgui -I$C/externals/stb -I$C/externals/pugixml/src -I$C/externals/tiny_obj_loader \
 -I$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])') -I$($PY -c 'import pybind11;print(pybind11.get_include())')"
TARGET="$OUT/redmax_py$EXT"
PROBE="$OUT/redmax_probe$EXT"
FLAGTAG="$OUT/obj/.flags_O3_NDEBUG_v2"
if [ ! -f "$FLAGTAG" ]; then rm -rf "$OUT/obj"; mkdir -p "$OUT/obj"; fi
if [ -f "$TARGET" ] && [ -f "$PROBE" ] && [ "$PROBE" -nt "$HERE/ref_probe.cpp" ] && [ "$PROBE" -nt "$HERE/build_ref.sh" ] && [ -z "${FORCE:-}" ]; then
  echo "build_ref: $TARGET and $PROBE already built"
else
  if [ ! -f "$OUT/obj/.done" ] || [ -n "${FORCE:-}" ]; then
    i=0
    for s in $SRCS; do i=$((i+1)); b="$(basename "$s" .cpp)"; echo "g++ $FLAGS -c $s -o $OUT/obj/${i}_$b.o"; done | xargs -P "$(nproc)" -I{} sh -c "{}"
    touch "$OUT/obj/.done" "$FLAGTAG"
  fi
  g++ -shared -o "$TARGET" "$OUT"/obj/*.o -ldl
  # the probe: our own read-only pybind module, linked with the same reference objects
  # (minus the reference's python_interface.o, which defines the other module)
  # The probe module alone links an INSTRUMENTED compile of Simulation.cpp: two counters (Newton iterations = calls of
  # func_with_derivatives, line-search evaluations = calls of func, DH/Simulation.cpp:1171,1189) inserted by sed into a
  # scratch copy under the git-ignored _ref directory.  redmax_py itself stays the unmodified reference.
  mkdir -p "$OUT/probe_src"
  sed -e 's|^\( *\)(this->\*func_with_derivatives)(x, g, H, false);|\1(this->*func_with_derivatives)(x, g, H, false); ++g_probe_newton_iters;|' \
      -e 's|^\( *\)(this->\*func)(x + alpha \* dx, g_new);|\1(this->*func)(x + alpha * dx, g_new); ++g_probe_ls_evals;|' \
      "$C/projects/redmax/Simulation.cpp" > "$OUT/probe_src/Simulation_counted.cpp"
  [ "$(grep -c 'g_probe_' "$OUT/probe_src/Simulation_counted.cpp")" = "2" ] || { echo "build_ref: instrumentation did not apply" >&2; exit 1; }
  echo 'extern int g_probe_newton_iters, g_probe_ls_evals;' > "$OUT/probe_src/probe_counters.h"
  g++ $FLAGS -include "$OUT/probe_src/probe_counters.h" -c "$OUT/probe_src/Simulation_counted.cpp" -o "$OUT/probe_src/Simulation_counted.o"
  g++ $FLAGS -include "$OUT/probe_src/probe_counters.h" -c "$HERE/ref_probe.cpp" -o "$OUT/probe_src/probe.o"
  g++ -shared -o "$PROBE" $(ls "$OUT"/obj/*.o | grep -v python_interface | grep -v '_Simulation\.o$') "$OUT/probe_src/Simulation_counted.o" "$OUT/probe_src/probe.o" -ldl
  echo "build_ref: built $TARGET and $PROBE"
fi
# Scene assets the reference binary needs at run time (XML + meshes are data, not source;
# they live only in the git-ignored _ref directory so the reference arm can run on the GPU box).
for scene in pusher dclaw_rotate tactile_insertion stable_grasp; do
  mkdir -p "$OUT/assets/$scene"
  cp -rf "$REF/envs/assets/$scene"/* "$OUT/assets/$scene/"
done
# BASELINE configs[0] (RollingBallExp test_sim_speed.py): reference-only plumbing case, tools/ref_config0.py
mkdir -p "$OUT/assets/tactile_pad"
cp -rf "$REF/assets/tactile_pad"/* "$OUT/assets/tactile_pad/"
# synthetic 32x13 variant named by BASELINE.json (same scene, denser marker grid)
sed 's/resolution="13 10"/resolution="32 13"/' "$OUT/assets/pusher/pusher.xml" > "$OUT/assets/pusher/pusher_32x13.xml"
# rolling-ball scene with a 40x40 marker grid (same dynamics, small golden fixture)
sed 's/resolution="200 200"/resolution="40 40"/' "$OUT/assets/tactile_pad/tactile_pad.xml" > "$OUT/assets/tactile_pad/tactile_pad_40x40.xml"
# synthetic variants named by BASELINE.json configs[3] / [4]: TactileInsertion with 2 x (20x20)
# pads (same scene, denser marker grids), DClaw with 3 x (8x6) = 48-marker pads (every 6th marker of the reference's
# 302-marker fingertip spec up to 48, laid on an 8x6 image grid)
sed 's/resolution="13 10"/resolution="20 20"/g' "$OUT/assets/tactile_insertion/tactile_insertion.xml" > "$OUT/assets/tactile_insertion/tactile_insertion_20x20.xml"
$PY - "$OUT/assets/dclaw_rotate" <<'PYX'
import sys, os
d = sys.argv[1]
lines = open(os.path.join(d, "tactile", "dclaw_fingertip_tactile.txt")).read().strip().splitlines()
n, rows = int(lines[0]), lines[1:]
pick = [rows[i] for i in range(0, n, 6)][:48]
out = ["48"]
for k, r in enumerate(pick):
    f = r.split('" "')
    f[1] = "%d %d" % (k // 6, k % 6)
    out.append('" "'.join(f))
open(os.path.join(d, "tactile", "dclaw_fingertip_tactile_8x6.txt"), "w").write("\n".join(out) + "\n")
x = open(os.path.join(d, "dclaw_torque_control.xml")).read().replace("tactile/dclaw_fingertip_tactile.txt", "tactile/dclaw_fingertip_tactile_8x6.txt")
open(os.path.join(d, "dclaw_torque_control_8x6.xml"), "w").write(x)
PYX
# The reference's own python callers of the path (R/envs/*.py, R/utils/*.py, R/algorithms/gd.py, the gd config), staged
# UNMODIFIED under the git-ignored _ref directory with the scene assets beside them where the env files look for them
# (envs/assets): tests run these files against the reference module AND against tactilesimulation_b200.redmax
# (sys.modules['redmax_py']) -- tests/test_reference_callers.py, tests/test_gpu_reference_callers.py.
PYD="$OUT/py"
mkdir -p "$PYD/envs" "$PYD/utils" "$PYD/algorithms" "$PYD/cfg"
cp -f "$REF"/envs/*.py "$PYD/envs/"
cp -f "$REF"/utils/*.py "$PYD/utils/"
cp -f "$REF"/algorithms/gd.py "$PYD/algorithms/"
cp -f "$REF"/examples/TactilePushExp/cfg/gd_tactile.yaml "$PYD/cfg/"
rm -rf "$PYD/envs/assets"
cp -r "$OUT/assets" "$PYD/envs/assets"
echo "build_ref: assets in $OUT/assets, reference callers in $PYD"
