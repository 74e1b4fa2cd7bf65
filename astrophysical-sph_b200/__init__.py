"""sph-b200: B200-native per-step SPH core (kNN, density+EOS, pressure/AV force, octree gravity, leapfrog)
behind the C ABI of libsph_b200.so.  See DESIGN.md."""
__version__ = "0.1.0"
