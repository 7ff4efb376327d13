"""MOT-format result files for the rows the fused tracker step returns -- the format ``save_results`` writes in the
reference driver (tracker/track.py:247-273), which stays the caller's when ``track.py`` runs unchanged; this helper serves
callers of ``TrackEngine`` / ``TrackingPipeline`` that never build ``STrack`` objects.

Engine row: ``[track_id, x, y, w, h, cls, score, slot]`` (float64), one array per frame and sequence.
"""
import os


def format_rows(frame_id, rows, data_type="mot17"):
    """Lines for one frame. 'mot17': ``frame,id,x,y,w,h,1.0,-1,-1,-1``; 'default': ``frame,id,x,y,w,h,cls`` (track.py:263-270)."""
    out = []
    for r in rows:
        tid, x, y, w, h, cls = int(r[0]), float(r[1]), float(r[2]), float(r[3]), float(r[4]), r[5]
        if data_type == "default":
            out.append("%d,%d,%.2f,%.2f,%.2f,%.2f,%d\n" % (frame_id, tid, x, y, w, h, int(cls)))
        elif data_type == "mot17":
            out.append("%d,%d,%.2f,%.2f,%.2f,%.2f,1.0,-1,-1,-1\n" % (frame_id, tid, x, y, w, h))
        else:
            raise ValueError("data_type must be 'default' or 'mot17'")
    return out


def write_sequence(path, frames, data_type="mot17", first_frame_id=1):
    """frames: iterable of per-frame row arrays of ONE sequence (frame ids count from 1, track.py:173).  Returns the path."""
    folder = os.path.dirname(path)
    if folder:
        os.makedirs(folder, exist_ok=True)
    n = 0
    with open(path, "w") as f:
        for k, rows in enumerate(frames):
            lines = format_rows(first_frame_id + k, rows, data_type)
            f.writelines(lines)
            n += len(lines)
    if n == 0:
        raise ValueError("no tracks to write (the reference asserts len(results))")
    return path
