"""Empty shim: baseline/evaluate_camera.py imports matplotlib at module scope."""
