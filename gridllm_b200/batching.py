"""Continuous batching on the host side: one runner thread per engine that steps every request the worker holds TOGETHER.

Reference side: a worker drops a second assignment while busy (client/src/services/WorkerClientService.ts:500-505) and the server
hands out MAX_CONCURRENT_JOBS_PER_WORKER = 1 (server/src/config/index.ts:31).  With the limit raised (NativeWorker
max_concurrent > 1), every generate* call of NativeInferenceService becomes a SEQUENCE of the engine (gl_seq_open), and this
runner drives gl_batch_step: sequences join between steps (their prompt is prefilled on admission), leave when they reach
num_predict / a stop token / are cancelled, and each token is handed to the request's callback the step it is produced -- the
same callback contract as gl_generate's (return True to cancel; job_cancellation, JobScheduler.ts:530-536).

The runner owns the engine while it works: every engine call is made under the service's engine lock, so embeddings and
un-batched calls interleave between steps, never inside one.
"""
from __future__ import annotations

import collections
import concurrent.futures
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import numpy as np

GL_ERR_NOMEM = -6


@dataclass
class _Job:
    ids: np.ndarray
    num_predict: int
    ignore_eos: bool
    kw: Dict[str, Any]
    on_token: Optional[Callable[[int, float, bytes], bool]]
    future: "concurrent.futures.Future"
    out_ids: List[int] = field(default_factory=list)
    out_lps: List[float] = field(default_factory=list)


class BatchRunner:
    def __init__(self, eng, lock: threading.Lock, max_batch: int, generation_cls):
        self.eng = eng
        self.lock = lock
        self.max_batch = int(max_batch)
        self._Generation = generation_cls
        self._pending: "collections.deque[_Job]" = collections.deque()
        self._jobs: Dict[int, _Job] = {}
        self._cv = threading.Condition()
        self._stop = False
        self.steps = 0                       # batched steps taken
        self.rows = 0                        # sum over steps of sequences in the step (rows / steps = mean batch size)
        self.max_rows = 0
        self._thread = threading.Thread(target=self._loop, name="gridllm-batch-runner", daemon=True)
        self._thread.start()

    # ---- request side -----------------------------------------------------------------------------------
    def submit(self, ids, num_predict: int, ignore_eos: bool, kw: Dict[str, Any], on_token=None) -> "concurrent.futures.Future":
        fut: "concurrent.futures.Future" = concurrent.futures.Future()
        job = _Job(np.ascontiguousarray(ids, dtype=np.int32), int(num_predict), bool(ignore_eos), dict(kw), on_token, fut)
        with self._cv:
            if self._stop:
                fut.set_exception(RuntimeError("engine is shut down"))
                return fut
            self._pending.append(job)
            self._cv.notify()
        return fut

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify()
        self._thread.join(timeout=30)

    # ---- runner thread ------------------------------------------------------------------------------------
    def _loop(self) -> None:
        while True:
            with self._cv:
                while not self._pending and not self._jobs and not self._stop:
                    self._cv.wait()
                if self._stop:
                    break
            try:
                self._admit()
                if self._jobs:
                    self._step()
            except Exception as ex:          # an engine failure ends every request it holds; the runner itself lives on
                self._fail_all(ex)
        self._fail_all(RuntimeError("engine is shut down"))

    def _fail_all(self, ex: Exception) -> None:
        with self._cv:
            jobs = list(self._jobs.items())
            self._jobs.clear()
            pend = list(self._pending) if self._stop else []
            if self._stop:
                self._pending.clear()
        for slot, job in jobs:
            try:
                with self.lock:
                    self.eng.seq_close(slot)
            except Exception:
                pass
            if not job.future.done():
                job.future.set_exception(ex)
        for job in pend:
            if not job.future.done():
                job.future.set_exception(ex)

    def _admit(self) -> None:
        """prefill waiting requests into free slots (between steps: sequences join and leave at step boundaries).  Everything
        that is waiting goes to the engine in ONE call: the prompts share packed prompt passes (gl_seq_open_many)."""
        room = self.max_batch - len(self._jobs)
        if room <= 0:
            return
        with self._cv:
            batch = [self._pending.popleft() for _ in range(min(room, len(self._pending)))]
        if not batch:
            return
        many = getattr(self.eng, "seq_open_many", None)
        if many is not None and len(batch) > 1:
            try:
                with self.lock:
                    slots = many([j.ids for j in batch], [dict(j.kw, num_predict=j.num_predict, ignore_eos=j.ignore_eos) for j in batch])
            except Exception as ex:
                if getattr(ex, "code", None) == GL_ERR_NOMEM and self._jobs:
                    with self._cv:                   # nothing fits right now: wait for a sequence to leave
                        self._pending.extendleft(reversed(batch))
                    return
                # nothing was opened (the engine validates every prompt before it opens any): a bad request among them, or not
                # even one fits -- the one-by-one path below finds out which, and fails only the requests that deserve it
                slots = None
            if slots is not None:
                back = []
                for j, slot in zip(batch, slots):
                    if slot >= 0:
                        self._jobs[slot] = j
                    else:
                        back.append(j)
                if back:
                    if self._jobs:
                        with self._cv:
                            self._pending.extendleft(reversed(back))
                        return
                    batch = back                     # nothing is running and these still do not fit: fail them one by one below
                else:
                    return
        for k, job in enumerate(batch):
            try:
                with self.lock:
                    slot = self.eng.seq_open(job.ids, num_predict=job.num_predict, ignore_eos=job.ignore_eos, **job.kw)
            except Exception as ex:
                if getattr(ex, "code", None) == GL_ERR_NOMEM and self._jobs:
                    with self._cv:                   # no slot / pages right now: wait for a sequence to leave
                        self._pending.extendleft(reversed(batch[k:]))
                    return
                job.future.set_exception(ex)         # can never fit (or a real error): this request fails, the others go on
                continue
            self._jobs[slot] = job

    def _step(self) -> None:
        with self.lock:
            res = self.eng.batch_step(cap=max(8, self.max_batch))
        n_rows = 0
        for slot, tok, lp, done in res:
            job = self._jobs.get(slot)
            if job is None:
                continue
            n_rows += 1
            cancel = False
            if tok >= 0:
                job.out_ids.append(int(tok))
                job.out_lps.append(float(lp))
                if job.on_token is not None:
                    try:
                        cancel = bool(job.on_token(int(tok), float(lp), self.eng.token_piece(int(tok))))
                    except Exception:
                        cancel = True
            if done or cancel:
                self._finish(slot, job, cancelled=cancel and not done)
        self.steps += 1
        self.rows += n_rows
        self.max_rows = max(self.max_rows, n_rows)

    def _finish(self, slot: int, job: _Job, cancelled: bool) -> None:
        try:
            with self.lock:
                st = self.eng.seq_stats(slot)
                self.eng.seq_close(slot)
        except Exception as ex:
            self._jobs.pop(slot, None)
            job.future.set_exception(ex)
            return
        self._jobs.pop(slot, None)
        if cancelled:
            st.done_reason = 2
        st.eval_count = len(job.out_ids)
        gen = self._Generation(np.asarray(job.out_ids, dtype=np.int32), np.asarray(job.out_lps, dtype=np.float32), st)
        job.future.set_result(gen)
