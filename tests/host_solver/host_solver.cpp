// TEST INFRASTRUCTURE: compiles the block-cooperative solver source (csrc/solve_core.cuh,
// solve_cascade.cuh) for the HOST with a one-thread team, so its numerics can be checked
// against OpenCV in the CPU test-suite (no GPU needed).  Never loaded by the product path.
#include "../../soccernet_calibration_sportlight_b200/csrc/solve_cascade.cuh"

extern "C" int host_camera_solve(const float* preds, const double* line_pts, const CalSolveParams* P, int B,
                                 CalCameraRecord* out) {
  using namespace cal::solve;
  Workspace* ws = new Workspace();
  const Team T{0, 1};
  for (int b = 0; b < B; ++b) {
    Frame fr;
    fr.pred = preds + (size_t)b * NKP * 3;
    fr.line_pts = line_pts ? line_pts + (size_t)b * NKP * 2 : nullptr;
    solve_frame(T, *ws, *P, fr, out + b);
  }
  delete ws;
  return 0;
}

extern "C" int host_workspace_bytes(void) { return (int)sizeof(cal::solve::Workspace); }
