from v2ce_toolbox_b200.scripts.LDATI import sample_voxel_statistical  # noqa: F401
