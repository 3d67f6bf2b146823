"""B200-native implementation of the ROMS nonlinear baroclinic time step (ROMS/Nonlinear/main3d.F chain).

Public surface:
  roms_trunk_mgh_b200.ocean.Tile    device-resident tile exposing the reference's routine names
  roms_trunk_mgh_b200.synth         analytical (synthetic) grids / initial conditions for UPWELLING, SEAMOUNT, BENCHMARK
  roms_trunk_mgh_b200._lib          ctypes binding of include/roms_b200.h
  roms_trunk_mgh_b200.build         nvcc build of lib/libroms_b200.so (sm_100a)
"""
__all__ = ["ocean", "synth", "_lib", "build"]
