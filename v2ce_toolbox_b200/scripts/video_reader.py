"""Drop-in for the part of /root/reference/scripts/video_reader.py the hot path uses
(v2ce.py:333-335,170): ``VideoReader(path, color_mode='GRAY')`` with ``frame_count``
(settable), ``fps``, ``width``, ``height`` and ``read_frames_at_indices``.

Host I/O only (cv2.VideoCapture); unlike the reference, which re-seeks before every frame
(video_reader.py:307), consecutive indices are decoded sequentially.
"""
import logging

import cv2
import numpy as np

logger = logging.getLogger(__name__)


class VideoReader:
    def __init__(self, path=None, color_mode='RGB', insets=(0, 0)):
        self.insets = insets
        self.color_mode = color_mode
        self.vidcap = None
        self._next = None
        self._frame_count = None
        self.path = path

    @property
    def path(self):
        return self._path

    @path.setter
    def path(self, path):
        self.close()
        self._path = path
        if path is not None:
            self.vidcap = cv2.VideoCapture(str(path))
            if not self.vidcap.isOpened():
                raise IOError(f'cannot open video {path}')
            self._frame_count = int(self.vidcap.get(cv2.CAP_PROP_FRAME_COUNT))
            self._next = 0

    def close(self):
        if self.vidcap is not None:
            self.vidcap.release()
            self.vidcap = None

    @property
    def fps(self):
        return self.vidcap.get(cv2.CAP_PROP_FPS)

    @property
    def frame_count(self):
        return self._frame_count

    @frame_count.setter
    def frame_count(self, value):
        self._frame_count = value

    @property
    def width(self):
        return int(self.vidcap.get(cv2.CAP_PROP_FRAME_WIDTH))

    @property
    def height(self):
        return int(self.vidcap.get(cv2.CAP_PROP_FRAME_HEIGHT))

    def _postprocess(self, frame):
        if self.color_mode in ('GRAY', 'GREY'):
            frame = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        elif self.color_mode == 'RGB':
            frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
        if self.insets[0] > 0:
            p = int(frame.shape[1] * self.insets[0])
            frame = frame[:, p:-p]
        if self.insets[1] > 0:
            q = int(frame.shape[1] * self.insets[1])
            frame = frame[q:-q, :]
        return frame

    def read_frame_at_index(self, frame_idx):
        frame_idx = int(frame_idx)
        if frame_idx != self._next:
            self.vidcap.set(cv2.CAP_PROP_POS_FRAMES, frame_idx)
        ret, frame = self.vidcap.read()
        self._next = max(frame_idx, 0) + 1
        if not ret or frame is None:
            logger.error('Error: Failed to retrieve frame %d from movie %s' % (frame_idx, self.path))
            self._next = None
            return None
        return self._postprocess(frame)

    def read_frames_at_indices(self, frame_idxs):
        assert len(frame_idxs) > 0
        frames = [f for f in (self.read_frame_at_index(i) for i in frame_idxs) if f is not None]
        if not frames:
            logger.info('No frames read from movie %s' % self.path)
            return None
        return np.stack(frames, axis=0)

    def read_all_frames(self):
        return self.read_frames_at_indices(range(self.frame_count))
