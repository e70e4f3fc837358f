"""Harness stand-in for ``matplotlib`` (not installed offline): run_nerf.py and run_nerf_helpers.py import
``matplotlib.pyplot`` at module level but only the dead debug branches draw.  TEST HARNESS ONLY."""
