// fpbh_internal.h -- shared declarations inside libfpb_host.so
#pragma once
#include <string>

#include "../../include/fpb_host.h"

int fpbh_fail(const char *fmt, ...);

// conformal map (cmap.cpp)
namespace cmap {
float cspanf(float value, float begin, float end);
void cnllxy(const float *m, float xlat, float xlong, float &xi, float &eta);
void cll2xy(const float *m, float xlat, float xlong, float &x, float &y);
void cxy2ll(const float *m, float x, float y, float &xlat, float &xlong);
void cc2gll(const float *m, float xlat, float xlong, float ue, float vn, float &ug, float &vg);
void stlmbr(float *m, float tnglat, float xlong);
void stcm2p(float *m, float x1, float y1, float xlat1, float xlong1, float x2, float y2,
            float xlat2, float xlong2);
} // namespace cmap
