/* oracle/_ref only: the one translation unit that instantiates stb_image for
 * src/imagetexture.cpp:20 / src/normalmap.cpp:20 (nanovg normally provides it). */
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image.h>
