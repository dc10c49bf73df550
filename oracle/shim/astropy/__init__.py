"""Stub (oracle only): see oracle/shim/README.md."""
