/* Shim for the reference build (oracle/_ref only): src/render.cpp:28 and src/main_euler.cpp:20
 * include <nori/gui.h>, which drags in nanogui/OpenGL.  Neither file uses anything from it except
 * what nori/render.h already declares, so the headless build substitutes this header. */
#pragma once
#include <nori/common.h>
#include <nori/render.h>
