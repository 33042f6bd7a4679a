/* Shim for the reference build (oracle/_ref only): include/nori/common.h:35 pulls
 * <ImathPlatform.h> from OpenEXR just to get the math constants; <math.h> is all it needs. */
#pragma once
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
