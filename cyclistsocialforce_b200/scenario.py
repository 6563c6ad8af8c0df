"""``Scenario`` runner with the reference's interface (reference
src/cyclistsocialforce/scenario.py:53-265): a fixed-step loop around a user step
function.  Animation / video write-out and SUMO co-simulation are out of scope
(presentation and host-side I/O); ``t_r`` real-time throttling is kept.
"""
from __future__ import annotations

from datetime import timedelta
from time import sleep, time


class Scenario:
    def __init__(self, step_func, t_0=0, t_s=0.01, t_r=0.01, animate=False, axes=None, verbose=True,
                 t_snapshots=(), write_animation=False, dir_animation_out=None, fname_animation_out=None,
                 tempdir_animation=None, keep_animation_frames=False, realtime=True):
        if animate or write_animation:
            raise NotImplementedError("matplotlib animation is outside the accelerated stepping path")
        self.t = t_0
        self.t_s = t_s
        self.t_r = t_r
        self.t_0 = t_0
        self.t_wall = time()
        self.i = 0
        self.animate = False
        self.ax = axes
        self.verbose = verbose
        self.step_func = step_func
        self.realtime = realtime

    def run(self, t_end):
        """reference :96-113 (without the blocking input() prompt of verbose mode)."""
        t_start = time()
        self._run_silent(t_start, t_end)
        if self.verbose:
            print(f"\nSimulation finished after {str(timedelta(seconds=time() - t_start))[:-3]}")

    def _run_silent(self, t_start, t_end):
        self.i_end = int(t_end / self.t_s)
        len_prev_msg = 0
        while self.i < self.i_end:
            t = time()
            self._step()
            len_prev_msg = self._wait(t, t_start, self.i_end, len_prev_msg)

    def _step(self):
        self.step_func()
        self.i += 1
        self.t += self.t_s

    def _wait(self, t, t_start, i_end, len_prev_msg):
        dt = time() - t
        t_sleep = max(0, self.t_r - dt) if self.realtime else 0.0
        msg = ""
        if self.verbose:
            sim_time = str(timedelta(seconds=self.t))[:11]
            wall_time = str(timedelta(seconds=(time() - t_start)))[:11]
            msg = (f"Running step {self.i}/{i_end}, Sim. time {sim_time}, Wall time {wall_time}, "
                   f"Wall freq. {int(1 / max(dt + t_sleep, 1e-9))} Hz ")
            msg += " " * max(len_prev_msg - len(msg), 0)
            print("\r" + msg, end="")
        if t_sleep > 0:
            sleep(t_sleep)
        return len(msg)

    def reset(self):
        self.i = 0
        self.t = self.t_0
