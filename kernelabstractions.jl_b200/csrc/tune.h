// tune.h -- run-time tuning knobs (b200_tune_set / b200_tune_get in the C ABI).
// Not part of the reference-facing surface: lets one gpurun call sweep kernel variants
// (grid sizing, unroll, implementation choice) without rebuilding.  Defaults are the measured best.
#pragma once
namespace b200 {
int tune_get(const char *key, int dflt);
void tune_set(const char *key, int value);
}  // namespace b200
