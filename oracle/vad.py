"""CPU restatement (numpy / plain Python) of the reference's frame -> event post-processing.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED: tests/golden/make_golden_vad.py ran the unmodified
/root/reference/utils/vad.py and utilities.frame_prediction_to_event_prediction on seeded inputs and stored
their outputs (tests/golden/vad_golden.json); tests/test_oracle_vad.py replays them.

Follows:
  activity_detection                       /root/reference/utils/vad.py:11-41
  find_bgn_fin_pairs                       /root/reference/utils/vad.py:44-66
  activity_detection_with_second_thres     /root/reference/utils/vad.py:69-91
  smooth                                   /root/reference/utils/vad.py:94-119
  remove_salt_noise                        /root/reference/utils/vad.py:122-134
  frame_prediction_to_event_prediction     /root/reference/utils/utilities.py:70-123

The reference's index arithmetic has asymmetries that are part of its observable behaviour and are kept:
the first pair starts AT the first active frame while later pairs start one frame AFTER their run's first
frame; every pair but the last ends one frame PAST its run, the last ends ON its run's last frame; ``smooth``
closes the final merged pair with the LAST pair's end (not the maximum).  Written here as one streaming pass
per series (the formulation the CUDA kernel uses) rather than as list transformations.
"""
import numpy as np


class _Smooth(object):
    """Streaming form of vad.smooth(): merge a pair into the previous one when bgn - previous_fin <= n."""

    def __init__(self, n, sink):
        self.n, self.sink, self.has = n, sink, False

    def push(self, bgn, fin):
        if not self.has:
            self.mem_bgn, self.has = bgn, True
        elif bgn - self.pre_fin > self.n:
            self.sink(self.mem_bgn, self.pre_fin)
            self.mem_bgn = bgn
        self.pre_fin = fin

    def finish(self):
        if self.has:
            self.sink(self.mem_bgn, self.pre_fin)


def activity_detection(x, thres, low_thres=None, n_smooth=1, n_salt=0):
    """Returns the list of [bgn, fin] frame-index pairs of vad.activity_detection.  Thresholds are compared in
    the array's own precision (fp32 for the model's outputs), as numpy does for a Python-float scalar.
    Raises IndexError where the reference does (a non-first run that begins on the last frame makes
    vad.py:78 read x[len(x)])."""
    x = np.asarray(x)
    T = len(x)
    thres = x.dtype.type(thres)
    out = []

    def salt(bgn, fin):
        if fin - bgn > n_salt:
            out.append([int(bgn), int(fin)])

    final = _Smooth(n_smooth, salt)
    if low_thres is None:
        stage = final
    else:
        low = x.dtype.type(low_thres)
        inner = _Smooth(1, final.push)

        class _Extend(object):
            @staticmethod
            def push(bgn, fin):
                while bgn != -1:
                    if x[bgn] < low:            # IndexError for bgn == T, exactly like the reference
                        break
                    bgn -= 1
                while fin != T:
                    if x[fin] < low:
                        break
                    fin += 1
                inner.push(bgn + 1, fin)

            @staticmethod
            def finish():
                inner.finish()
                final.finish()
        stage = _Extend

    # runs of frames above the high threshold -> the reference's asymmetric [bgn, fin] pairs
    run = 0
    s = e = -1
    for t in range(T):
        if x[t] > thres:
            if e == t - 1 and s >= 0:
                e = t
            else:
                if s >= 0:
                    stage.push(s if run == 0 else s + 1, e + 1)
                    run += 1
                s = e = t
    if s >= 0:
        stage.push(s if run == 0 else s + 1, e)
    if low_thres is None:
        final.finish()
    else:
        stage.finish()
    return out


def _per_class(v, classes_num):
    return list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] * classes_num


def frame_prediction_to_event_prediction(output_dict, sed_params_dict, frames_per_second, labels):
    """utilities.py:70-123 with config.frames_per_second / config.labels passed explicitly."""
    audios_num, frames_num, classes_num = output_dict['framewise_output'].shape
    p = {k: _per_class(sed_params_dict[k], classes_num) for k in
         ('audio_tagging_threshold', 'sed_high_threshold', 'sed_low_threshold', 'n_smooth', 'n_salt')}
    events = []
    for n in range(audios_num):
        for k in range(classes_num):
            if output_dict['clipwise_output'][n, k] > p['audio_tagging_threshold'][k]:
                pairs = activity_detection(output_dict['framewise_output'][n, :, k], p['sed_high_threshold'][k],
                                           p['sed_low_threshold'][k], p['n_smooth'][k], p['n_salt'][k])
                for bgn, fin in pairs:
                    events.append({'filename': output_dict['audio_name'][n],
                                   'onset': bgn / float(frames_per_second),
                                   'offset': fin / float(frames_per_second),
                                   'event_label': labels[k]})
    return events
