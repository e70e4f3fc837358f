"""Harness stand-in for ``pyvista`` (not installed offline): load/load_scannet.py:8 imports it at module level for a
mesh viewer that the Blender / LLFF paths never touch.  TEST HARNESS ONLY."""
