"""Empty shim (see __init__)."""
