/* b2o_f64.h -- force-included (-include) by the DOUBLE-PRECISION build of the CPU oracle, libb2o64.so.
 *
 * TEST INFRASTRUCTURE ONLY (see b2o_world.h).
 *
 * Purpose: BASELINE.json's north star asks for "body poses within a stated fp32 tolerance after 240 substeps".
 * pybullet (which computes in double) cannot run here, so the tolerance is stated against the same algorithm carried
 * out in double precision: every `float` of the oracle's sources and of the shared leaf headers becomes `double`,
 * and the single-precision polynomial sin/cos/atan2 are replaced by libm's.  What stays float: the C-ABI structures
 * of include/b2s.h (B2SParams, B2SSceneDesc), which are included below before the redefinition, so the Python side
 * builds them exactly as for the fp32 libraries.  Arrays that cross b2o_* entry points are double in this build
 * (oracle/b2o.py: OracleWorld(..., f64=True)).
 */
#ifndef B2O_F64_H_
#define B2O_F64_H_

/* everything whose declarations must keep the real `float` comes first (include guards keep it that way) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../include/b2s.h"

typedef float abi_float;     /* the float of the C-ABI structures, which stays single precision */

#define B2S_F64 1
#define float double
#define sqrtf sqrt
#define fabsf fabs
#define fminf fmin
#define fmaxf fmax
#define floorf floor
#define rintf rint
#define fmaf fma

#endif
