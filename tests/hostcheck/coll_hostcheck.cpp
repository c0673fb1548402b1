// coll_hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles the per-cell functions of loki_b200/csrc/lk_coll.cuh (the bodies
// the CUDA kernels of lk_coll.cu run) for the HOST and walks them over a box with the kernels' own index maps, so that
// the arithmetic can be compared with the oracle in the container that has no GPU (tests/test_cpu_coll_host.py).
// Nothing under loki_b200/ loads this; the library has no CPU path.
#include "../../loki_b200/csrc/lk_coll.cuh"

using namespace lkcoll;

extern "C" void hc_fields(double* ivx, double* ivy, double* vth, const double* u, const int* n, int ng, const double* dx,
                          const double* vel) {
  const int nd0 = n[0] + 2 * ng, nd1 = n[1] + 2 * ng, nd2 = n[2] + 2 * ng, nd3 = n[3] + 2 * ng;
  const i64 pl = (i64)nd0 * nd1, pv = (i64)nd2 * nd3;
  for (i64 c2 = 0; c2 < pl; ++c2)
    fields_point(u + c2, pl, ng, n[2], n[3], nd2, vel, pv, dx[2] * dx[3], ivx[c2], ivy[c2], vth[c2]);
}

extern "C" void hc_append(double* rhs, const double* f, const int* n, int ng, int order, const double* dx, const double* vel,
                          const double* ivx, const double* ivy, const double* vth, const double* vlo, const double* vhi,
                          const double* rlo, const double* rhi, double vfloor, double nu_coef, int conservative) {
  if (!conservative && order != 4) return;
  const int nd0 = n[0] + 2 * ng, nd1 = n[1] + 2 * ng, nd2 = n[2] + 2 * ng, nd3 = n[3] + 2 * ng;
  const i64 pl = (i64)nd0 * nd1;
  Params P;
  const int vr = order == 4 ? 3 : 4;
  for (int k = 0; k < 2; ++k) {
    P.range_lo[k] = rlo[k];
    P.range_hi[k] = rhi[k];
    P.vmin[k] = vlo[k] + vr * dx[2 + k];
    P.vmax[k] = vhi[k] - vr * dx[2 + k];
  }
  P.vfloor = vfloor;
  P.nu_coef = nu_coef;
  P.dvx = dx[2];
  P.dvy = dx[3];
  const i64 total = (i64)n[0] * n[1] * n[2] * n[3];
  for (i64 t = 0; t < total; ++t) {
    const int i1 = (int)(t % n[0]) + ng;
    i64 r = t / n[0];
    const int i2 = (int)(r % n[1]) + ng;
    r /= n[1];
    const int i3 = (int)(r % n[2]) + ng, i4 = (int)(r / n[2]) + ng;
    const i64 c2 = i1 + (i64)nd0 * i2;
    const double cf = order == 4 ? collision_cell<4>(f, pl, nd2, nd3, c2, i3, i4, vel, ivx, ivy, vth, P, conservative)
                                 : collision_cell<6>(f, pl, nd2, nd3, c2, i3, i4, vel, ivx, ivy, vth, P, conservative);
    const i64 c = c2 + pl * (i3 + (i64)nd2 * i4);
    rhs[c] = rhs[c] + cf;
  }
}
