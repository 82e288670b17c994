"""``Logger`` facade (dronesim/utils/Logger.py): the reference's rollout log format, fed from the GPU.

The data model is the reference's (Logger.py:22-157): ``timestamps[num_drones, T]``,
``states[num_drones, state_length, T]``, ``controls[num_drones, control_length, T]``, filled by
``log(drone, timestamp, state, control)`` with the same growth rules, saved by ``save()`` as one
``np.savez`` archive with keys ``timestamps / states / controls`` (the reference names the file ``.npy``).

Batched addition: ``attach(env, vehicles)`` asks the CUDA core to record the aviary state vector of the chosen
vehicles after every step ON THE DEVICE (``ds_log_attach``: one tiny kernel per step writing straight into the
``states[drone][state][sample]`` layout), and ``collect()`` pulls the whole log in one copy - so a million-vehicle
rollout can still produce the reference's per-drone traces for a handful of vehicles without a per-step
host round trip.  Plotting (matplotlib, Logger.py:161-426) is out of scope.
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np


class Logger(object):
    """A class for logging (and, in the reference, visualization)."""

    def __init__(self, logging_freq_hz: int, state_length: int = 20, control_length: int = 12, num_drones: int = 1,
                 duration_sec: int = 0):
        self.state_length = state_length
        self.control_length = control_length
        self.LOGGING_FREQ_HZ = logging_freq_hz
        self.NUM_DRONES = num_drones
        self.PREALLOCATED_ARRAYS = False if duration_sec == 0 else True
        self.counters = np.zeros(num_drones)
        self.timestamps = np.zeros((num_drones, duration_sec * self.LOGGING_FREQ_HZ))
        self.states = np.zeros((num_drones, self.state_length, duration_sec * self.LOGGING_FREQ_HZ))
        self.controls = np.zeros((num_drones, self.control_length, duration_sec * self.LOGGING_FREQ_HZ))
        self._core = None

    # ------------------------------------------------------------------ reference API
    def log(self, drone: int, timestamp, state, control=np.zeros(12)):
        """Logger.log (Logger.py:87-139), same validation message and array-growth rules."""
        if (drone < 0 or drone >= self.NUM_DRONES or timestamp < 0 or len(state) != self.state_length
                or len(control) != self.control_length):
            print(f" State Length : {self.state_length}, Control Length : {self.control_length}")
            print("[ERROR] in Logger.log(), invalid data")
        current_counter = int(self.counters[drone])
        #### Add rows to the matrices if a counter exceeds their size (:115-125)
        if current_counter >= self.timestamps.shape[1]:
            self.timestamps = np.concatenate((self.timestamps, np.zeros((self.NUM_DRONES, 1))), axis=1)
            self.states = np.concatenate((self.states, np.zeros((self.NUM_DRONES, self.state_length, 1))), axis=2)
            self.controls = np.concatenate((self.controls, np.zeros((self.NUM_DRONES, self.control_length, 1))), axis=2)
        #### Advance a counter if the matrices have overgrown it (:127-130)
        elif not self.PREALLOCATED_ARRAYS and self.timestamps.shape[1] > current_counter:
            current_counter = self.timestamps.shape[1] - 1
        self.timestamps[drone, current_counter] = timestamp
        self.states[drone, :, current_counter] = state
        self.controls[drone, :, current_counter] = control
        self.counters[drone] = current_counter + 1

    def save(self, file_path=None, file_name=None):
        """Logger.save (Logger.py:143-157): one np.savez archive, keys timestamps / states / controls."""
        if file_path is None:
            file_path = os.path.join(os.getcwd(), "files", "logs") + os.sep
            os.makedirs(file_path, exist_ok=True)
        if file_name is None:
            file_name = "save-flight-" + datetime.now().strftime("%m.%d.%Y_%H.%M.%S")
        path = file_path + file_name + ".npy"
        with open(path, "wb") as out_file:
            np.savez(out_file, timestamps=self.timestamps, states=self.states, controls=self.controls)
        return path

    # ------------------------------------------------------------------ device capture
    def attach(self, env, vehicles=None, env_index: int = 0, capacity: int = None):
        """Record on the device.  ``env``: a batched aviary facade or a ``SwarmCore``; ``vehicles``: global vehicle
        ids, default = the NUM_DRONES drones of environment ``env_index``."""
        core = getattr(env, "_core", env)
        if vehicles is None:
            vehicles = env_index * core.D + np.arange(self.NUM_DRONES)
        vehicles = np.asarray(vehicles, dtype=np.int32).reshape(-1)
        if vehicles.shape[0] != self.NUM_DRONES:
            print("[ERROR] in Logger.attach(), %d vehicles for a %d-drone logger" % (vehicles.shape[0], self.NUM_DRONES))
            raise ValueError("vehicles")
        if capacity is None:
            capacity = self.timestamps.shape[1] if self.PREALLOCATED_ARRAYS else 4096
        core.log_attach(vehicles, int(capacity))
        self._core = core

    def collect(self, controls=None):
        """Pull the device log into ``timestamps / states`` (``controls`` [num_drones, control_length, T] optional)."""
        if self._core is None:
            print("[ERROR] in Logger.collect(), no device log attached")
            raise RuntimeError("attach() first")
        ts, st = self._core.log_read()
        T = ts.shape[0]
        self.timestamps = np.tile(ts[None, :], (self.NUM_DRONES, 1))
        self.states = st[:, : self.state_length, :].copy()
        self.controls = np.zeros((self.NUM_DRONES, self.control_length, T)) if controls is None else np.asarray(controls, float)
        self.counters = np.full(self.NUM_DRONES, float(T))
        return T
