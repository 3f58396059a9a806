// invmgmt_jit.cuh -- run-time specialisation of the serial multi-echelon env's fused rollout (see invmgmt_jit.cu)
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "invmgmt_jit_args.cuh"

struct InvJitSpec {  // host-side copy of everything the generator bakes into the kernel
    int n, T, backlog;
    int L[ORGYM_INV_MAX_STAGES];
    long long c[ORGYM_INV_MAX_STAGES], I0[ORGYM_INV_MAX_STAGES];
    double up[ORGYM_INV_MAX_STAGES + 1], uc[ORGYM_INV_MAX_STAGES + 1], kc[ORGYM_INV_MAX_STAGES + 1],
        hc[ORGYM_INV_MAX_STAGES + 1];
    std::vector<double> disc;  // alpha**t
    int log2k, base;           // alias table geometry of the demand distribution
    long long xbound;          // bound on |on-hand|, |backlog|, |sales| of any stage in any period (0 = unknown)
};
// true when every float64 operation of a period's profit is exact for |integers| <= xbound (see invmgmt_jit.cu)
bool inv_jit_profit_is_exact(const InvJitSpec& S);
#define INV_JIT_THREADS 128
// configurations the generator covers (everything else runs the ahead-of-time kernel)
bool inv_jit_eligible(const InvJitSpec& S);
// complete translation unit with two kernels: inv_jit_rollout_bs (base-stock, integer levels), inv_jit_rollout_rnd
std::string inv_jit_source(const InvJitSpec& S);
