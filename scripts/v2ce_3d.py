from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d  # noqa: F401
