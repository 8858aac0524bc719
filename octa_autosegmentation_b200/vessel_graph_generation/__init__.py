"""The reference's object API for the growth path (SURVEY 8b "Python API used by main": generate_vessel_graph.py:24-56,
example_custom_vessel_simulation.ipynb) over the GPU engine -- same module names, class names, attributes and method names:

    from octa_autosegmentation_b200.vessel_graph_generation.greenhouse import Greenhouse
    from octa_autosegmentation_b200.vessel_graph_generation.forest import Forest
    from octa_autosegmentation_b200.vessel_graph_generation import tree2img

    greenhouse = Greenhouse(config["Greenhouse"], seed=0)
    art = Forest(config["Forest"], greenhouse.d, greenhouse.r, greenhouse.simspace, nerve_center=greenhouse.nerve_center, ...)
    ven = Forest(config["Forest"], greenhouse.d, greenhouse.r, greenhouse.simspace, arterial=False, ...)
    greenhouse.set_forests(art, ven)
    greenhouse.develop_forest()            # ONE GPU growth run (growth.GrowContext); the forests are filled from its edge tables
    for tree in art.get_trees():
        for node in tree.get_tree_iterator(exclude_root=True, only_active=False):
            node.position, node.get_proximal_node().position, node.radius

What differs from the reference, by construction: the simulation does not run inside these Python objects, so the random
streams are the engine's -- `Greenhouse(cfg, seed=s)` gives exactly the graph the unmodified reference grows after
`random.seed(s); np.random.seed(s)` issued right before `Greenhouse(...)` (without `seed`, one is drawn from Python's `random`,
so a seeded script stays reproducible); the trees exist once `develop_forest()` has run, not after `Forest(...)`; a run without a
venous forest (`set_forests(art)`) is not offered by the engine.  Many samples at once: `pipeline.Pipeline` / the CLIs."""
