"""Inference-side callers of the front-end that avoid redundant extraction.

Both reproduce, bit for bit or to rounding, what the reference obtains by calling its extractor
again and again on transformed or overlapping waveforms:

* ACS test-time augmentation (`BaseModelModule.post_processing`,
  /root/reference/src/models/components/model_module.py:272-284) runs the extractor on 16
  sign-flipped / channel-swapped copies of every batch.  Log-mel is invariant to a sign flip and
  the intensity vector I_j = Re(conj(X_0) X_j) just changes sign / slot, so all 16 feature maps
  follow from ONE extraction.
* Overlapped-chunk inference (`test_hoplen_sec: 0.5` with 10-s chunks,
  configs/data/*/evalMA.yaml:11-12; chunking by `segment_index`, src/utils/data_utilities.py:6-64)
  featurises every audio second 20 times.  Frames whose analysis window lies inside a chunk are
  identical to the same frames of the whole recording; only the 3 + 3 frames that touch a chunk
  edge see that chunk's reflect padding and are recomputed from two short excerpts.
"""
import torch

# model_module.py:273-275 -- loop order of the reference: 8 sign triples x 2 channel orders
ACS_SIGNS = [[1, 1, 1], [-1, 1, 1], [1, -1, 1], [-1, -1, 1],
             [1, 1, -1], [-1, 1, -1], [1, -1, -1], [-1, -1, -1]]
ACS_TRANS = [((0, 1, 2), (1, 2, 3)), ((1, 0, 2), (3, 2, 1))]


def acs_waveform_variants(batch_sample):
    """The 16 waveform batches the reference builds (model_module.py:276-282), in its order."""
    for sign in ACS_SIGNS:
        for _trans_y, trans_x in ACS_TRANS:
            sign_x, sign_y, sign_z = sign
            s_x, s_y, s_z = trans_x
            yield torch.stack((batch_sample[:, 0], sign_y * batch_sample[:, s_x],
                               sign_z * batch_sample[:, s_y], sign_x * batch_sample[:, s_z]), axis=1)


def acs_feature_variants(feat):
    """feat = LogmelIV_Extractor(x) for the un-augmented batch x (B, 4, L) -> (B, 7, T, M).
    Yields the 16 feature maps the reference gets from `self.standardize(variant)`, in its order,
    without touching the waveform again: channel c' of a variant is +-channel src[c'] of x, so
    log-mel c' = log-mel src[c'] and IV_j' = sign_j * IV_src[j]."""
    if feat.ndim != 4 or feat.shape[1] != 7:
        raise ValueError('expected (B, 7, T, M) log-mel+IV features of a 4-channel FOA batch')
    for sign in ACS_SIGNS:
        for _trans_y, trans_x in ACS_TRANS:
            sign_x, sign_y, sign_z = sign
            s_x, s_y, s_z = trans_x
            out = torch.empty_like(feat)
            out[:, 0] = feat[:, 0]
            out[:, 1] = feat[:, s_x]
            out[:, 2] = feat[:, s_y]
            out[:, 3] = feat[:, s_z]
            out[:, 4] = sign_y * feat[:, 3 + s_x]
            out[:, 5] = sign_z * feat[:, 3 + s_y]
            out[:, 6] = sign_x * feat[:, 3 + s_z]
            yield out


def segment_index(x_len, chunklen, hoplen, last_frame_always_paddding=False):
    """(begin, end) and (pad_before, pad_after) per chunk: src/utils/data_utilities.py:6-64."""
    if x_len < chunklen:
        return [(0, x_len)], [(0, chunklen - x_len)]
    n_frames = 1 + (x_len - chunklen) // hoplen
    idx = [(n * hoplen, n * hoplen + chunklen) for n in range(n_frames)]
    pad = [(0, 0)] * n_frames
    if (n_frames - 1) * hoplen + chunklen == x_len:
        return idx, pad
    if last_frame_always_paddding or x_len - n_frames * hoplen >= chunklen // 2:
        idx.append((n_frames * hoplen, x_len))
        pad.append((0, chunklen - (x_len - n_frames * hoplen)))
    else:
        idx.append((x_len - chunklen, x_len))
        pad.append((0, 0))
    return idx, pad


def extract_overlapped(extractor, x, chunklen, hoplen, last_frame_always_paddding=False):
    """Features of every inference chunk of one recording x (C, L) on the extractor's device:
    returns ((n_chunks, C+3, 1 + chunklen//hop, M), chunk index list), equal bit for bit to
    `extractor(stack(zero-padded chunks))`, with the recording transformed once."""
    if x.ndim != 2:
        raise ValueError('x must be (channels, samples) of one recording')
    C, L = x.shape
    hop, n_fft = extractor.hop, extractor.n_fft
    idx, pad = segment_index(L, chunklen, hoplen, last_frame_always_paddding)
    T = 1 + chunklen // hop
    half = n_fft // 2
    n_edge = -(-half // hop)                           # frames whose window crosses a chunk edge (3)
    t_hi = (chunklen - half) // hop                    # last frame whose window ends inside the chunk
    fast = [i for i, ((b, e), (pb, pa)) in enumerate(zip(idx, pad))
            if pb == 0 and pa == 0 and b % hop == 0 and chunklen % hop == 0 and chunklen > 4 * n_fft]
    slow = [i for i in range(len(idx)) if i not in fast]
    full = None
    out = None
    if fast:
        full = extractor(x.unsqueeze(0))[0]            # (C+3, 1 + L//hop, M): the recording, once
        out = torch.empty((len(idx), full.shape[0], T, full.shape[2]), dtype=full.dtype, device=full.device)
        head_len = (n_edge - 1) * hop + half + hop     # covers the windows of frames 0..n_edge-1
        n_tail = T - 1 - t_hi                          # frames after t_hi
        tail_frames = n_tail + n_edge + 1              # excerpt long enough that its first kept frame is interior
        tail_len = (tail_frames - 1) * hop
        heads = torch.stack([x[:, idx[i][0]: idx[i][0] + head_len] for i in fast])
        tails = torch.stack([x[:, idx[i][1] - tail_len: idx[i][1]] for i in fast])
        fh = extractor(heads)                          # (n, C+3, 1 + head_len//hop, M)
        ft = extractor(tails)                          # (n, C+3, tail_frames, M)
        for j, i in enumerate(fast):
            f0 = idx[i][0] // hop
            out[i, :, n_edge:t_hi + 1] = full[:, f0 + n_edge: f0 + t_hi + 1]
            out[i, :, :n_edge] = fh[j, :, :n_edge]
            out[i, :, t_hi + 1:] = ft[j, :, tail_frames - n_tail:]
    if slow:
        chunks = torch.stack([torch.nn.functional.pad(x[:, idx[i][0]: idx[i][1]], (pad[i][0], pad[i][1])) for i in slow])
        fs = extractor(chunks)
        if out is None:
            out = torch.empty((len(idx),) + tuple(fs.shape[1:]), dtype=fs.dtype, device=fs.device)
        for j, i in enumerate(slow):
            out[i] = fs[j]
    return out, idx
