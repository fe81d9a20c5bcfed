"""TEST INFRASTRUCTURE (oracle shim). Minimal stand-in for the un-vendored `pytorch3d` package so the
reference's own files import; only the rotation conversions the hot path calls are provided."""
