"""Import-time stub for matplotlib (absent from this image).  TEST INFRASTRUCTURE ONLY.
The reference imports pyplot/collections/cm at module scope (greenhouse.py:2, tree2img.py:10);
nothing on the seeded growth / voxelize path calls into them.  rasterize_forest (Agg) is NOT
runnable through this stub -> 2-D raster parity is unpinned in-container."""
