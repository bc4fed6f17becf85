"""B200-native calibration hot path (keypoint / line HRNet forward, heat-map decode,
camera solve) behind the reference's Python API.  See DESIGN.md."""
__version__ = "0.1.0"
