"""Reference-side plug-ins that make a fullrmc Engine draw the counter-based random numbers of
:mod:`fullrmc_b200.rng` -- the host half of the contract behind ``frmc_run_generated``.

fullrmc lets a user supply the group selector (``Engine.set_group_selector``) and every group's move generator
(``Group.set_move_generator``); the only other draw of a step is the acceptance test's ``generate_random_float``, a
module-level name of ``fullrmc.Engine`` (Engine.py:33, :3311).  ``install`` uses exactly those three seams -- no
reference source is touched -- so that ``Engine.run`` walks, bit for bit, the trajectory the device walks when
``DeviceStore.run_generated`` is given the same seed and first counter (tests/gen_golden_generated.py runs the
unmodified Engine that way; tests/test_generated_runs.py replays on the device).

The classes need fullrmc's base classes, so they are built by a factory from the imported package.
"""
import numpy as np

from . import rng


class CounterStream(object):
    """step counter + the words of the current step; select_index opens a step (Engine.py:3168 is the first draw)"""

    def __init__(self, seed, first_counter=0):
        self.seed = int(seed)
        self.next_counter = int(first_counter)
        self.words = None

    def open_step(self):
        self.words = rng.step_words(self.seed, self.next_counter)
        self.next_counter += 1
        return self.words

    def acceptance(self):
        return float(rng.acceptance_number(self.words))


def make_classes(fullrmc):
    from fullrmc.Globals import INT_TYPE, FLOAT_TYPE
    from fullrmc.Selectors.RandomSelectors import RandomSelector
    from fullrmc.Generators.Translations import TranslationGenerator

    class CounterRandomSelector(RandomSelector):
        """RandomSelector whose index is word 0 of the step: (w * numberOfGroups) >> 32"""

        def __init__(self, engine, stream):
            super(CounterRandomSelector, self).__init__(engine)
            self._stream = stream

        def select_index(self):
            w = self._stream.open_step()
            return INT_TYPE(rng.group_index(w[0], len(self.engine.groups)))

    class CounterTranslationGenerator(TranslationGenerator):
        """TranslationGenerator whose vector comes from the step's words (same amplitude semantics)"""

        def __init__(self, stream, group=None, amplitude=0.2):
            super(CounterTranslationGenerator, self).__init__(group=group, amplitude=amplitude)
            self._stream = stream

        def transform_coordinates(self, coordinates, argument=None):
            lo, hi = self.amplitude
            return coordinates + rng.translation_vector(self._stream.words, FLOAT_TYPE(lo), FLOAT_TYPE(hi))

    return CounterRandomSelector, CounterTranslationGenerator


def install(fullrmc, engine, seed, first_counter=0, amplitude=0.2):
    """Put the counter-based selector, generators and acceptance number into `engine`; returns the stream."""
    import fullrmc.Engine as engine_module
    Selector, Generator = make_classes(fullrmc)
    stream = CounterStream(seed, first_counter)
    engine.set_group_selector(Selector(engine, stream))
    for g in engine.groups:
        g.set_move_generator(Generator(stream, amplitude=amplitude))
    engine_module.generate_random_float = stream.acceptance
    return stream
