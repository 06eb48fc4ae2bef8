/* Synthetic input generator for the drfe tests and bench.py — NOT part of the product ABI.
 * Deterministic procedural RGB-D frame (textured Manhattan corridor / room): gray u8 (w*h) and depth f32
 * (w*h, metres * depth_unit_scale), seed selects the camera pose.  scene: 0 = corridor, 1 = room,
 * 2 = room + pillars.  Returns 0, or -1 on a bad argument. */
#ifndef DRFE_SYNTH_H_
#define DRFE_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int drfe_synth_frame(int width, int height, int scene, uint32_t seed, float depth_unit_scale,
                     uint8_t* gray, float* depth, float* fx, float* fy, float* cx, float* cy);
#ifdef __cplusplus
}
#endif
#endif
