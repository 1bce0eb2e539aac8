"""Oracle: ``librosa.pyin`` (librosa 0.10.x, numpy 1.x semantics) restated in NumPy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**: librosa is an
un-vendored, unpinned dependency of the reference (``requirements.txt``); the reference
has no test or golden vector for it.  Call site being restated:
``roar/collections/tts/data/dataset.py:696-703``::

    librosa.pyin(audio.numpy(), fmin=pitch_fmin, fmax=pitch_fmax,
                 frame_length=win_length, sr=sample_rate, fill_na=0.0)

so ``hop_length = frame_length // 4``, ``win_length = frame_length // 2``, ``center=True``,
``pad_mode="constant"``, 100 thresholds, beta(2, 18), boltzmann 2, resolution 0.1,
max_transition_rate 35.92, switch_prob 0.01, no_trough_prob 0.01.

Published algorithm followed (librosa 0.10.x): ``core/pitch.py`` ``pyin``,
``_cumulative_mean_normalized_difference``, ``_parabolic_interpolation``, ``__pyin_helper``;
``sequence.py`` ``viterbi``/``_viterbi``, ``transition_local``, ``transition_loop``;
``util/utils.py`` ``localmin``.  numpy-1.x detail kept on purpose: ``np.fft.rfft`` of the
float32 frames runs in float64 while the energy cumsum stays float32.
"""
import numpy as np
import numba
import scipy.signal
import scipy.stats


# ----------------------------------------------------------------------------- framing
def frame_signal(y, frame_length, hop_length):
    """``np.pad(constant)`` + ``librosa.util.frame`` -> float32 ``[frame_length, T]``."""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, (frame_length // 2, frame_length // 2), mode="constant")
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    return yp[idx]


def periods(sr, fmin, fmax, frame_length, win_length):
    min_period = int(np.floor(sr / fmax))
    max_period = min(int(np.ceil(sr / fmin)), frame_length - win_length - 1)
    return min_period, max_period


# ----------------------------------------------------------------------------- CMND
def difference_function(y_frames, frame_length, win_length):
    """YIN difference function d[tau], tau = 0 .. frame_length-win_length-1, float64
    ``[n_lags_all, T]``; autocorrelation via float64 FFT, energy via float32 cumsum."""
    yf64 = y_frames.astype(np.float64)
    a = np.fft.rfft(yf64, frame_length, axis=0)
    b = np.fft.rfft(yf64[win_length:0:-1, :], frame_length, axis=0)
    acf = np.fft.irfft(a * b, frame_length, axis=0)[win_length:, :]
    acf[np.abs(acf) < 1e-6] = 0
    energy = np.cumsum(y_frames ** 2, axis=0)  # float32, sequential
    energy = energy[win_length:, :] - energy[:-win_length, :]
    energy[np.abs(energy) < 1e-6] = 0
    return energy[:1, :] + energy - 2 * acf  # (f32 + f32) - f64


def cmnd(y_frames, frame_length, win_length, min_period, max_period):
    """Cumulative-mean-normalised difference, float64 ``[max_period-min_period+1, T]``."""
    d = difference_function(y_frames, frame_length, win_length)
    num = d[min_period : max_period + 1, :]
    tau = np.arange(1, max_period + 1)[:, None]
    cmean = np.cumsum(d[1 : max_period + 1, :], axis=0) / tau
    den = cmean[min_period - 1 : max_period, :]
    return num / (den + np.finfo(den.dtype).tiny)


def parabolic_shifts(x):
    """``_parabolic_interpolation`` along axis 0."""
    shifts = np.zeros_like(x)
    a = x[2:] + x[:-2] - 2 * x[1:-1]
    b = (x[2:] - x[:-2]) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        s = -b / a
    s[np.abs(b) >= np.abs(a)] = 0
    shifts[1:-1] = s
    return shifts


def localmin(x):
    """``librosa.util.localmin`` on a 1-D array."""
    m = np.zeros(x.shape, dtype=bool)
    m[1:-1] = (x[1:-1] < x[:-2]) & (x[1:-1] <= x[2:])
    m[-1] = x[-1] < x[-2]
    return m


# ----------------------------------------------------------------------------- observation
def beta_threshold_prior(n_thresholds=100, beta_parameters=(2, 18)):
    thresholds = np.linspace(0, 1, n_thresholds + 1)
    beta_cdf = scipy.stats.beta.cdf(thresholds, beta_parameters[0], beta_parameters[1])
    return thresholds, np.diff(beta_cdf)


def n_pitch_bins_for(fmin, fmax, resolution=0.1):
    n_bins_per_semitone = int(np.ceil(1.0 / resolution))
    return int(np.floor(12 * n_bins_per_semitone * np.log2(fmax / fmin))) + 1, n_bins_per_semitone


def frame_trough_probs(yin_frame, thresholds, beta_probs, boltzmann_parameter, no_trough_prob):
    """Per-frame part of ``__pyin_helper``: -> (trough_index, probs) or (None, None)."""
    is_trough = localmin(yin_frame)
    is_trough[0] = yin_frame[0] < yin_frame[1]
    (trough_index,) = np.nonzero(is_trough)
    if len(trough_index) == 0:
        return None, None
    heights = yin_frame[trough_index]
    below = np.less.outer(heights, thresholds[1:])
    positions = np.cumsum(below, axis=0) - 1
    n_troughs = np.count_nonzero(below, axis=0)
    with np.errstate(all="ignore"):
        prior = scipy.stats.boltzmann.pmf(positions, boltzmann_parameter, n_troughs)
    prior[~below] = 0
    probs = prior.dot(beta_probs)
    gmin = np.argmin(heights)
    n_below_min = np.count_nonzero(~below[gmin, :])
    probs[gmin] += no_trough_prob * np.sum(beta_probs[:n_below_min])
    return trough_index, probs


def observation_probs(yin_frames, shifts, sr, fmin, min_period, n_pitch_bins, n_bins_per_semitone,
                      thresholds, beta_probs, boltzmann_parameter=2, no_trough_prob=0.01):
    """``__pyin_helper`` -> (obs float64 ``[2*npb, T]``, voiced_prob ``[T]``)."""
    yin_probs = np.zeros_like(yin_frames)
    for i in range(yin_frames.shape[1]):
        ti, probs = frame_trough_probs(yin_frames[:, i], thresholds, beta_probs,
                                       boltzmann_parameter, no_trough_prob)
        if ti is not None:
            yin_probs[ti, i] = probs
    yin_period, frame_index = np.nonzero(yin_probs)
    period_candidates = min_period + yin_period
    period_candidates = period_candidates + shifts[yin_period, frame_index]
    f0_candidates = sr / period_candidates
    bin_index = 12 * n_bins_per_semitone * np.log2(f0_candidates / fmin)
    bin_index = np.clip(np.round(bin_index), 0, n_pitch_bins).astype(int)
    obs = np.zeros((2 * n_pitch_bins, yin_frames.shape[1]))
    # duplicates: NumPy fancy assignment keeps the last write (np.nonzero is row-major,
    # so within a frame the larger lag wins).  Made explicit here.
    for b, f, p in zip(bin_index, frame_index, yin_probs[yin_period, frame_index]):
        obs[b, f] = p
    voiced_prob = np.clip(np.sum(obs[:n_pitch_bins, :], axis=0, keepdims=True), 0, 1)
    obs[n_pitch_bins:, :] = (1 - voiced_prob) / n_pitch_bins
    return obs, voiced_prob[0]


# ----------------------------------------------------------------------------- HMM
def transition_local(n_states, width):
    """``librosa.sequence.transition_local(n_states, width, window="triangle", wrap=False)``."""
    transition = np.zeros((n_states, n_states), dtype=np.float64)
    win = scipy.signal.get_window("triangle", width, fftbins=False)
    lpad = (n_states - width) // 2
    for i in range(n_states):
        row = np.zeros(n_states)
        row[lpad : lpad + width] = win  # util.pad_center
        row = np.roll(row, n_states // 2 + i + 1)
        row[min(n_states, i + width // 2 + 1) :] = 0
        row[: max(0, i - width // 2)] = 0
        transition[i] = row
    transition /= transition.sum(axis=1, keepdims=True)
    return transition


def transition_loop(n_states, prob):
    transition = np.empty((n_states, n_states), dtype=np.float64)
    for i in range(n_states):
        transition[i] = (1.0 - prob) / (n_states - 1)
        transition[i, i] = prob
    return transition


def hmm_tables(sr, hop_length, n_pitch_bins, n_bins_per_semitone,
               max_transition_rate=35.92, switch_prob=0.01):
    """-> (transition ``[2npb, 2npb]``, p_init ``[2npb]``, transition_width)."""
    max_semitones_per_frame = round(max_transition_rate * 12 * hop_length / sr)
    transition_width = max_semitones_per_frame * n_bins_per_semitone + 1
    local = transition_local(n_pitch_bins, transition_width)
    t_switch = transition_loop(2, 1 - switch_prob)
    transition = np.kron(t_switch, local)
    p_init = np.zeros(2 * n_pitch_bins)
    p_init[n_pitch_bins:] = 1 / n_pitch_bins
    return transition, p_init, transition_width


@numba.njit(cache=True)
def _viterbi_dense(log_prob, log_trans, log_p_init):
    """``librosa.sequence._viterbi``: dense DP, first-index argmax."""
    n_steps, n_states = log_prob.shape
    state = np.zeros(n_steps, dtype=np.uint16)
    value = np.zeros((n_steps, n_states), dtype=np.float64)
    ptr = np.zeros((n_steps, n_states), dtype=np.uint16)
    value[0] = log_prob[0] + log_p_init
    for t in range(1, n_steps):
        trans_out = value[t - 1] + log_trans.T
        for j in range(n_states):
            ptr[t, j] = np.argmax(trans_out[j])
            value[t, j] = log_prob[t, j] + trans_out[j, ptr[t][j]]
    state[-1] = np.argmax(value[-1])
    for t in range(n_steps - 2, -1, -1):
        state[t] = ptr[t + 1, state[t + 1]]
    return state


@numba.njit(cache=True)
def _viterbi_banded(log_prob, log_trans, log_p_init, npb, hw):
    """Same DP, same result (first-index argmax over all predecessors), but the
    out-of-band predecessors -- whose log-transition is the constant ``log(tiny)`` --
    are covered by prefix/suffix maxima.  Test speed-up only; checked against
    ``_viterbi_dense`` in tests/test_oracle_pyin.py."""
    n_steps, n_states = log_prob.shape
    state = np.zeros(n_steps, dtype=np.uint16)
    ptr = np.zeros((n_steps, n_states), dtype=np.uint16)
    prev = log_prob[0] + log_p_init
    cur = np.empty(n_states)
    pmax = np.empty(n_states)
    parg = np.empty(n_states, dtype=np.int64)
    smax = np.empty(n_states)
    sarg = np.empty(n_states, dtype=np.int64)
    lt0 = log_trans[0, n_states // 2 - 1]  # an out-of-band entry: log(0 + tiny)
    for t in range(1, n_steps):
        for blk in range(2):
            o = blk * npb
            pmax[o] = prev[o]
            parg[o] = o
            for i in range(1, npb):
                if prev[o + i] > pmax[o + i - 1]:
                    pmax[o + i] = prev[o + i]
                    parg[o + i] = o + i
                else:
                    pmax[o + i] = pmax[o + i - 1]
                    parg[o + i] = parg[o + i - 1]
            smax[o + npb - 1] = prev[o + npb - 1]
            sarg[o + npb - 1] = o + npb - 1
            for i in range(npb - 2, -1, -1):
                if prev[o + i] >= smax[o + i + 1]:
                    smax[o + i] = prev[o + i]
                    sarg[o + i] = o + i
                else:
                    smax[o + i] = smax[o + i + 1]
                    sarg[o + i] = sarg[o + i + 1]
        for j in range(n_states):
            jb = j % npb
            best = -np.inf
            arg = 0
            for blk in range(2):
                o = blk * npb
                lo = jb - hw
                hi = jb + hw
                if lo > 0:
                    s = pmax[o + lo - 1] + lt0
                    if s > best:
                        best = s
                        arg = parg[o + lo - 1]
                for i in range(max(lo, 0), min(hi, npb - 1) + 1):
                    s = prev[o + i] + log_trans[o + i, j]
                    if s > best:
                        best = s
                        arg = o + i
                if hi < npb - 1:
                    s = smax[o + hi + 1] + lt0
                    if s > best:
                        best = s
                        arg = sarg[o + hi + 1]
            ptr[t, j] = arg
            cur[j] = log_prob[t, j] + best
        for j in range(n_states):
            prev[j] = cur[j]
    state[-1] = np.argmax(prev)
    for t in range(n_steps - 2, -1, -1):
        state[t] = ptr[t + 1, state[t + 1]]
    return state


def viterbi(obs, transition, p_init, banded=None):
    """``librosa.sequence.viterbi`` (log-domain with ``tiny`` guard).  ``banded=(npb, hw)``
    selects the equivalent fast DP."""
    eps = np.finfo(obs.dtype).tiny
    log_trans = np.log(transition + eps)
    log_prob = np.ascontiguousarray(np.log(obs.T + eps))
    log_p_init = np.log(p_init + eps)
    if banded is None:
        return _viterbi_dense(log_prob, log_trans, log_p_init)
    return _viterbi_banded(log_prob, log_trans, log_p_init, banded[0], banded[1])


# ----------------------------------------------------------------------------- top level
def pyin(y, fmin, fmax, sr=22050, frame_length=2048, win_length=None, hop_length=None,
         n_thresholds=100, beta_parameters=(2, 18), boltzmann_parameter=2, resolution=0.1,
         max_transition_rate=35.92, switch_prob=0.01, no_trough_prob=0.01, fill_na=np.nan,
         dense_viterbi=False, return_internals=False):
    """-> (f0 float64 ``[T]``, voiced_flag bool ``[T]``, voiced_prob float64 ``[T]``)."""
    if win_length is None:
        win_length = frame_length // 2
    if hop_length is None:
        hop_length = frame_length // 4
    y_frames = frame_signal(y, frame_length, hop_length)
    min_period, max_period = periods(sr, fmin, fmax, frame_length, win_length)
    yin_frames = cmnd(y_frames, frame_length, win_length, min_period, max_period)
    shifts = parabolic_shifts(yin_frames)
    thresholds, beta_probs = beta_threshold_prior(n_thresholds, beta_parameters)
    npb, nbps = n_pitch_bins_for(fmin, fmax, resolution)
    obs, voiced_prob = observation_probs(yin_frames, shifts, sr, fmin, min_period, npb, nbps,
                                         thresholds, beta_probs, boltzmann_parameter, no_trough_prob)
    transition, p_init, tw = hmm_tables(sr, hop_length, npb, nbps, max_transition_rate, switch_prob)
    states = viterbi(obs, transition, p_init, banded=None if dense_viterbi else (npb, tw // 2))
    freqs = fmin * 2 ** (np.arange(npb) / (12 * nbps))
    f0 = freqs[states % npb]
    voiced_flag = states < npb
    if fill_na is not None:
        f0[~voiced_flag] = fill_na
    if return_internals:
        return f0, voiced_flag, voiced_prob, dict(yin_frames=yin_frames, shifts=shifts, obs=obs,
                                                  states=states, npb=npb, tw=tw,
                                                  min_period=min_period, max_period=max_period)
    return f0, voiced_flag, voiced_prob
