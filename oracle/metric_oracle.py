"""ts_diff_metric -- CPU restatement of /root/reference/train/scripts/stage2/stage2_metrics.py:22-88.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Vectorised per cell instead of per event, same arithmetic: the smallest
|t_pred - t_gt| over the predicted events of the same polarity within the (clamped) search window, 1e6 when there is
none, capped at 1e6/fps/10*3 with an overflow count; the mean over the ground-truth events."""
import numpy as np


def ts_diff_metric_oracle(event_gt, event_pred, search_range=0, fps=30, width=346, height=260):
    gt_p = np.asarray(event_gt['polarity']).astype(np.int64).copy()
    gt_p[gt_p == -1] = 0
    cells = {}
    for t, x, y, p in zip(event_pred['timestamp'], event_pred['x'], event_pred['y'], event_pred['polarity']):
        cells.setdefault((int(x), int(y), int(p)), []).append(int(t))
    cells = {k: np.array(v, dtype=np.int64) for k, v in cells.items()}
    cap = 1e6 / fps / 10 * 3
    total, overflow = 0, 0
    for t, x, y, p in zip(event_gt['timestamp'], event_gt['x'], event_gt['y'], gt_p):
        diff = 1e6
        for a in range(max(int(x) - search_range, 0), min(int(x) + search_range + 1, width)):
            for b in range(max(int(y) - search_range, 0), min(int(y) + search_range + 1, height)):
                c = cells.get((a, b, int(p)))
                if c is not None:
                    diff = min(diff, np.min(np.abs(c - int(t))))
        if diff > cap:
            diff = cap
            overflow += 1
        total += diff
    return np.array([total / len(event_gt), overflow])
