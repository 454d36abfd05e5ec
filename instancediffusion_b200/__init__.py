"""instancediffusion_b200 -- B200-native (sm_100a) kernels + drop-in host mirror for the
InstanceDiffusion sampling hot path.  See DESIGN.md."""

__version__ = "0.1.0"
