"""ORACLE (test infrastructure) — independent numpy twins of the Appendix-A primitives.

The torch restatement in ``oracle/model.py`` is the oracle proper; these loop-level numpy versions
re-derive the same arithmetic from the published semantics so that the two can
be checked against each other (validation (ii)).  Small sizes only.
"""
import numpy as np


def conv1d_same(x, W):
    """A.3: TF Conv1D SAME, stride 1, cross-correlation; pad_left=(k-1)//2.  x [B,T,Ci], W [k,Ci,Co]."""
    B, T, Ci = x.shape
    k, _, Co = W.shape
    pl = (k - 1) // 2
    y = np.zeros((B, T, Co), dtype=np.float64)
    for b in range(B):
        for t in range(T):
            for j in range(k):
                s = t + j - pl
                if 0 <= s < T:
                    y[b, t] += x[b, s] @ W[j]
    return y


def maxpool2_same(x):
    """A.3: MaxPooling1D(2, stride 1, SAME): right pad with -inf."""
    y = x.copy()
    y[:, :-1] = np.maximum(x[:, :-1], x[:, 1:])
    return y


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, c, h, W, b, forget_bias=1.0):
    """A.5: TF LSTMCell gate order i, j, f, o."""
    z = np.concatenate([x, h], -1) @ W + b
    H = c.shape[-1]
    i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
    c_new = sigmoid(f + forget_bias) * c + sigmoid(i) * np.tanh(j)
    return c_new, sigmoid(o) * np.tanh(c_new)


def softmax_masked(e, length):
    m = e[:length].max()
    p = np.zeros_like(e)
    p[:length] = np.exp(e[:length] - m)
    return p / p.sum()


def forward_attention_step(query, prev_align, prev_alpha, keys, length, Wq, conv_w, conv_b, Wf, v, ba, u=0.5):
    """forward_attention.py:88-122 for ONE batch row.  keys [T,A]; conv_w [k,1,F]; returns (a, alpha)."""
    T, A = keys.shape
    k = conv_w.shape[0]
    pl = (k - 1) // 2
    q = query @ Wq
    f = np.zeros((T, conv_w.shape[2]))
    for t in range(T):
        for j in range(k):
            s = t + j - pl
            if 0 <= s < T:
                f[t] += prev_align[s] * conv_w[j, 0]
    f += conv_b
    e = (v * np.tanh(keys + q[None, :] + f @ Wf + ba)).sum(-1)
    a = softmax_masked(e, length)
    shifted = np.concatenate([[0.0], prev_alpha[:-1]])
    alpha = ((1 - u) * prev_alpha + u * shifted + 1e-7) * a
    return a, alpha / alpha.sum()
