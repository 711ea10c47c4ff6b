"""Imported automatically at interpreter start-up when this directory is on PYTHONPATH: see eonerf_shim.py."""
import eonerf_shim

eonerf_shim.install()
