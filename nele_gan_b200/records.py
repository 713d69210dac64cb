"""Score record format between the labelling step and the discriminator's data loader
(SURVEY.md section 8f, rank 2).

The reference serialises every labelled utterance as the string
``"siib,haspi,estoi,pesq,visqol,path"`` (``List_concat_5scores`` + ``List_concat``,
audio_util.py:367-389) and the discriminator dataset splits it again
(``Discriminator_train_dataset.__getitem__``, dataloader.py:54-84: fields 0-2 -> ``True_score``
float32[3], fields 3-4 -> ``True_score_Qua`` float32[2], field 5 -> WAV path).  The functions
here keep that wire format byte for byte -- ``str(float)`` of the float64 scores, comma
separated -- so the reference's loader reads what the engine produces, and add the direct
tensor form for a loader that stays in memory."""
import numpy as np


def List_concat(score, enhanced_list):                       # audio_util.py:367-371
    return [str(score[i]) + ',' + enhanced_list[i] for i in range(len(score))]


def List_concat_score(score, score2):                        # audio_util.py:373-377
    return [str(score[i]) + ',' + str(score2[i]) for i in range(len(score))]


def List_concat_3scores(score1, score2, score3):             # audio_util.py:379-383
    return [str(score1[i]) + ',' + str(score2[i]) + ',' + str(score3[i]) for i in range(len(score1))]


def List_concat_5scores(score1, score2, score3, score4, score5):   # audio_util.py:385-389
    return [','.join(str(s[i]) for s in (score1, score2, score3, score4, score5)) for i in range(len(score1))]


def round_records(scores, enhanced_list, pesq=None, visqol=None):
    """Records of one sampling round (train_nele.py:320-328) from the engine's ``[n, 3]`` score
    matrix {SIIB, HASPI, ESTOI}.  PESQ and ViSQOL are external binaries in the reference
    (intel.py:142-160, audio_util.py:323-365) and stay so: pass their mapped scores, or leave them
    out and the two quality fields are written as 0.0."""
    scores = np.asarray(scores, dtype=np.float64)
    n = scores.shape[0]
    if len(enhanced_list) != n:
        raise ValueError("scores and enhanced_list differ in length")
    pesq = [0.0] * n if pesq is None else list(pesq)
    visqol = [0.0] * n if visqol is None else list(visqol)
    five = List_concat_5scores([float(v) for v in scores[:, 0]], [float(v) for v in scores[:, 1]],
                               [float(v) for v in scores[:, 2]], pesq, visqol)
    return List_concat(five, list(enhanced_list))


def parse_record(record):
    """What dataloader.py:56-79 extracts: ``(True_score float32[3], True_score_Qua float32[2], path)``."""
    f = record.split(',')
    return (np.asarray([float(f[0]), float(f[1]), float(f[2])], dtype=np.float32),
            np.asarray([float(f[3]), float(f[4])], dtype=np.float32), f[5])


def records_to_tensors(records):
    """All records of a round as ``(True_score [n, 3], True_score_Qua [n, 2], paths)`` float32 arrays."""
    parsed = [parse_record(r) for r in records]
    return (np.stack([p[0] for p in parsed]) if parsed else np.zeros((0, 3), np.float32),
            np.stack([p[1] for p in parsed]) if parsed else np.zeros((0, 2), np.float32),
            [p[2] for p in parsed])
