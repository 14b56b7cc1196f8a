"""AirAssembly programs written for the tests (inputs of the front-end, genstark_b200/assembly.py) with plain
Python control implementations of what they compute."""
from genstark_b200.air import P128, prng_sha256

# MiMC over p128: x' = x^3 + k[step mod 64]; the start value arrives through the seed vector
MIMC_SOURCE = '''
(module
    (field prime 340282366920938463463374607393113505793)
    (const $cube scalar 3)
    (function $round
        (result vector 1)
        (param $x vector 1) (param $key scalar)
        (add (exp (load.param $x) (load.const $cube)) (load.param $key)))
    (export mimc
        (registers 1) (constraints 1) (steps STEPS)
        (static
            (cycle (prng sha256 0x4d694d43 64)))     # round keys
        (init
            (param $start vector 1)
            (load.param $start))
        (transition
            (call $round (load.trace 0) (get (load.static 0) 0)))
        (evaluation
            (sub
                (load.trace 1)
                (call $round (load.trace 0) (get (load.static 0) 0))))))
'''

# A 4-register sponge: per block a secret two-element start state, per block `words` public message words, each
# absorbed by 16 steps of  state' = MIX * (state + key[step])^3  (+ the next word in register 0 at the end of its 16 steps).
#   statics 0,1  block start state (secret, one value per block, held one step early: shift -1)
#           2    message words (public, children of 0, 16 steps each, shift -1)
#           3,4  masks: last step of a block / last step of a word
#           5..8 round keys
SPONGE_MIX = [[2, 1, 1, 3], [3, 2, 1, 1], [1, 3, 2, 1], [1, 1, 3, 2]]
SPONGE_SOURCE = '''
(module
    (field prime 340282366920938463463374607393113505793)
    (const $mix matrix (2 1 1 3) (3 2 1 1) (1 3 2 1) (1 1 3 2))
    (function $permute
        (result vector 4)
        (param $state vector 4) (param $keys vector 4)
        (prod
            (load.const $mix)
            (exp (add (load.param $state) (load.param $keys)) (scalar 3))))
    (function $step
        (result vector 4)
        (param $state vector 4) (param $k vector 9)
        (local $next vector 4)
        (store.local $next
            (add
                (call $permute (load.param $state) (slice (load.param $k) 5 8))
                (mul
                    (vector (get (load.param $k) 2) (scalar 0) (scalar 0) (scalar 0))
                    (get (load.param $k) 4))))
        (add
            (mul
                (vector (slice (load.param $k) 0 1) (get (load.param $k) 2) (scalar 0))
                (get (load.param $k) 3))
            (mul
                (load.local $next)
                (sub (scalar 1) (get (load.param $k) 3)))))
    (export sponge
        (registers 4) (constraints 4) (steps 16)
        (static
            (input secret (shift -1))
            (input secret (peerof 0) (shift -1))
            (input public (childof 0) (steps 16) (shift -1))
            (mask (input 0))
            (mask (input 2))
            (cycle (prng sha256 0x73706f6e676531 16))
            (cycle (prng sha256 0x73706f6e676532 16))
            (cycle (prng sha256 0x73706f6e676533 16))
            (cycle (prng sha256 0x73706f6e676534 16)))
        (init
            (vector (slice (load.static 0) 0 1) (get (load.static 0) 2) (scalar 0)))
        (transition
            (call $step (load.trace 0) (load.static 0)))
        (evaluation
            (sub (load.trace 1) (call $step (load.trace 0) (load.static 0))))))
'''


def sponge_inputs(blocks: int, words: int):
    a = [(1000 + 17 * b) % P128 for b in range(blocks)]
    c = [(P128 - 5 - b) % P128 for b in range(blocks)]
    msg = [[(b * 1315423911 + w * 2654435761 + 7) % P128 for w in range(words)] for b in range(blocks)]
    return [a, c, msg]


def sponge_control(inputs, blocks: int, words: int):
    """register-major trace computed the slow, obvious way"""
    p = P128
    keys = [prng_sha256(bytes.fromhex('73706f6e67653%d' % (i + 1)), 16, p) for i in range(4)]
    a, c, msg = inputs
    T = blocks * words * 16
    rows = []
    st = None
    for s in range(T):
        b, w, r = s // (words * 16), (s // 16) % words, s % 16
        if s % (words * 16) == 0:
            st = [a[b] % p, c[b] % p, msg[b][0] % p, 0]        # block start: (start state, first word, 0)
        rows.append(list(st))
        t = [pow((st[j] + keys[j][r]) % p, 3, p) for j in range(4)]
        st = [sum(SPONGE_MIX[i][j] * t[j] for j in range(4)) % p for i in range(4)]
        if r == 15 and w + 1 < words:
            st[0] = (st[0] + msg[b][w + 1]) % p                  # the next word enters register 0
    return [[rows[s][r] for s in range(T)] for r in range(4)]
