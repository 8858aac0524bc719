"""Minimal stand-in for matplotlib (absent from this image).  TEST INFRASTRUCTURE ONLY.

The reference imports pyplot/collections/cm at module scope (greenhouse.py:2, tree2img.py:10).  Growth and voxelization never
call into them.  rasterize_forest (tree2img.py:51-108) uses exactly this surface:
    plt.figure(figsize), figure.patch.set_facecolor('black'), plt.axes([0,0,1,1], frameon=False, xticks=[], yticks=[]),
    ax.invert_yaxis(), collections.LineCollection(edges, linewidths, colors='w', antialiaseds=True, capstyle='round'),
    ax.add_collection, figure.canvas.draw(), canvas.buffer_rgba(), canvas.get_width_height(), plt.close(figure)
which the shim implements by handing the collection's segments and line widths to oracle/agg_oracle.c, the restatement of
the Agg pipeline those calls run in real matplotlib (it reproduces all 500 label PNGs the reference ships bit for bit).  With
it the UNMODIFIED rasterize_forest runs in the build container (oracle/ref_harness.rasterize)."""
