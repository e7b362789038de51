// bp_pick.h — kernel pickers exported by the bp_inst.cu objects (one per precision x method).
#pragma once

namespace qd {

struct BpParams;
struct Bp7Params;
typedef void (*bp_kernel_fn)(const BpParams);
typedef void (*bp7_kernel_fn)(const Bp7Params);

#ifndef QD_BP7_MAXT
#define QD_BP7_MAXT 800
#endif
constexpr int BP7_MAXT = QD_BP7_MAXT;   // launch bound of the row-regular kernel (sets the register budget)

#define QD_DECL_PICK(tag)                                        \
    bp_kernel_fn pick_bp6_##tag(int vmax, int threads);          \
    bp7_kernel_fn pick_bp7_##tag(int rd, int vmax, int slots, int irregular, int cluster, int threads);
QD_DECL_PICK(f64_ms)
QD_DECL_PICK(f64_ps)
QD_DECL_PICK(f32_ms)
QD_DECL_PICK(f32_ps)
#undef QD_DECL_PICK

}  // namespace qd
