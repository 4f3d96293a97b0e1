from v2ce_toolbox_b200.scripts.video_reader import VideoReader  # noqa: F401
