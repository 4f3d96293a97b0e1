"""Module paths of the reference (``scripts.v2ce_3d``, ``scripts.LDATI``, ``scripts.video_reader``)
re-exported from v2ce_toolbox_b200.scripts so existing imports keep working."""
