"""TEST / BENCH INFRASTRUCTURE: a restatement of the reference server's dispatch rules -- the part the north_star says is
"reused unchanged" and that shards independent requests across the in-process GPU workers.  The real server is TypeScript +
Redis and cannot run in this image; this drives N in-process NativeWorkers through the same pub/sub messages.

Restated, with the reference lines each rule follows (/root/reference/server/src/services/):
  * queue: in-memory array, stable sort by priority high > medium > low on every pass      JobScheduler.ts:145-151
  * one pass per tick (1 000 ms in the reference, `tick_s` here; 0 = as fast as events allow) JobScheduler.ts:128-135
  * candidates: status == "online" AND currentJobs < MAX_CONCURRENT_JOBS_PER_WORKER AND the worker lists the model
                                                                                            WorkerRegistry.ts:397-419
  * choice: fewest currentJobs, then performance tier high > medium > low (missing tier = 2), stable
                                                                                            JobScheduler.ts:337-349
  * assignment: markWorkerBusy (+1; status "busy" once the count reaches the maximum), publish
    worker:<id>:job {type: job_assignment, job: {jobId, workerId, request, assignedAt, timeout}}
                                                                                            JobScheduler.ts:362-414, WorkerRegistry.ts:421-458
  * completion / failure: markWorkerAvailable (-1; "busy" -> "online" below the maximum)     JobScheduler.ts:434-461, WorkerRegistry.ts:460-462
  * worker:status_update: status taken verbatim, currentJobs = data.currentJobs || worker.currentJobs (a reported 0 never
    clears the count -- only job completion does)                                           WorkerRegistry.ts:325-332
  * worker:heartbeat: status = data.status || "online", currentJobs = data.currentJobs || 0  WorkerRegistry.ts:261-270
  * job_cancellation message to the holder of a job                                          JobScheduler.ts:530-536, 888-894
Not restated: Redis persistence, timeouts / orphan scan / retries (no fault injection here), the HTTP gateway."""
import asyncio
import json
import time
from datetime import datetime, timezone

PRIO = {"high": 3, "medium": 2, "low": 1}
TIER = {"high": 3, "medium": 2, "low": 1}


def _js_or(value, fallback):
    """JavaScript `value || fallback`: 0, "", null / undefined are falsy; an (empty) array is truthy"""
    if value is None or value == 0 or value == "" or value is False:
        if isinstance(value, list):
            return value
        return fallback
    return value


class SchedulerStandIn:
    def __init__(self, bus, max_jobs_per_worker=1, tick_s=0.0):
        self.bus = bus
        self.workers = {}                  # insertion-ordered, like the registry's Map
        self.queue = []
        self.results = {}
        self.assigned = {}                 # jobId -> workerId
        self.active = {}                   # jobId -> workerId while the job runs
        self.max_jobs = max_jobs_per_worker
        self.tick_s = float(tick_s)
        self.ticks = 0
        self.stream_chunks = {}            # jobId -> number of job:stream chunks seen (when watched)
        self.first_chunk_at = {}
        self.submitted_at = {}
        self.done_at = {}

    async def start(self):
        await self.bus.subscribe("worker:registered", self._on_registered)
        await self.bus.subscribe("worker:status_update", self._on_status)
        await self.bus.subscribe("worker:heartbeat", self._on_heartbeat)
        await self.bus.subscribe("job:completed", self._on_done)
        await self.bus.subscribe("job:failed", self._on_done)

    # ---- WorkerRegistry ---------------------------------------------------------------------------------
    async def _on_registered(self, msg):
        d = json.loads(msg)
        self.workers[d["workerId"]] = {"workerId": d["workerId"], "capabilities": d["capabilities"], "status": "online", "currentJobs": 0}

    async def _on_status(self, msg):
        d = json.loads(msg)
        w = self.workers.get(d["workerId"])
        if w:
            w["status"] = d["status"]
            w["currentJobs"] = _js_or(d.get("currentJobs"), w["currentJobs"])

    async def _on_heartbeat(self, msg):
        d = json.loads(msg)
        w = self.workers.get(d["workerId"])
        if w:
            w["status"] = _js_or(d.get("status"), "online")
            w["currentJobs"] = _js_or(d.get("currentJobs"), 0)

    def _job_count(self, w, inc):
        cur = w["currentJobs"]
        cur = 0 if isinstance(cur, list) else int(cur)                 # [] + 1 is "1" in JS; Math.max(0, "1") is 1
        w["currentJobs"] = max(0, cur + inc)
        if w["currentJobs"] >= self.max_jobs:
            w["status"] = "busy"
        elif w["status"] == "busy":
            w["status"] = "online"

    async def _on_done(self, msg):
        d = json.loads(msg)
        self.results[d["jobId"]] = d
        self.done_at[d["jobId"]] = time.perf_counter()
        wid = self.active.pop(d["jobId"], None)                        # handleJobCompleted looks the assignment up by job id
        w = self.workers.get(wid) if wid else None
        if w:
            self._job_count(w, -1)

    # ---- JobScheduler -----------------------------------------------------------------------------------
    def available(self, model):
        def cur(w):
            return 0 if isinstance(w["currentJobs"], list) else w["currentJobs"]
        return [w for w in self.workers.values() if w["status"] == "online" and cur(w) < self.max_jobs
                and any(m["name"] == model for m in w["capabilities"]["availableModels"])]

    def select(self, job):
        cand = self.available(job["model"])
        cand.sort(key=lambda w: ((0 if isinstance(w["currentJobs"], list) else w["currentJobs"]),
                                 -TIER.get(w["capabilities"].get("performanceTier"), 2)))      # stable, like Array.prototype.sort
        return cand[0] if cand else None

    def add_job(self, request):
        self.submitted_at[request["id"]] = time.perf_counter()
        self.queue.append(request)

    async def watch_stream(self, job_id):
        async def on_chunk(_msg, jid=job_id):
            self.stream_chunks[jid] = self.stream_chunks.get(jid, 0) + 1
            self.first_chunk_at.setdefault(jid, time.perf_counter())
        await self.bus.subscribe(f"job:stream:{job_id}", on_chunk)

    async def cancel_job(self, job_id):
        wid = self.active.get(job_id)
        if wid:
            await self.bus.publish(f"worker:{wid}:job", json.dumps({"type": "job_cancellation", "jobId": job_id}))

    async def tick(self):
        """one pass of processJobQueue: priority order, assign while workers are available"""
        self.ticks += 1
        self.queue.sort(key=lambda j: -PRIO.get(j.get("priority") or "medium", 2))
        launched, rest = [], []
        for job in self.queue:
            w = self.select(job)
            if w is None:
                rest.append(job)
                continue
            self._job_count(w, +1)                                      # markWorkerBusy
            self.assigned[job["id"]] = w["workerId"]
            self.active[job["id"]] = w["workerId"]
            msg = json.dumps({"type": "job_assignment", "job": {"jobId": job["id"], "workerId": w["workerId"], "request": job,
                                                                "assignedAt": datetime.now(timezone.utc).isoformat(),
                                                                "timeout": job.get("timeout") or 300000}})
            launched.append(asyncio.ensure_future(self.bus.publish(f"worker:{w['workerId']}:job", msg)))
        self.queue = rest
        return launched

    async def run_until_empty(self):
        """tick until every queued job has a result.  tick_s = 0: passes are driven by completions; tick_s > 0: one pass per
        tick_s seconds of wall clock, the reference's dispatch granularity (1 s)."""
        pending = []
        n_jobs = len(self.queue) + len(self.active)
        want = set(j["id"] for j in self.queue) | set(self.active)
        next_tick = time.perf_counter()
        while self.queue or pending or any(j not in self.results for j in want):
            now = time.perf_counter()
            if self.tick_s <= 0 or now >= next_tick:
                pending += await self.tick()
                next_tick = now + self.tick_s
            pending = [p for p in pending if not p.done()]
            if self.tick_s > 0:
                await asyncio.sleep(min(0.002, max(0.0, next_tick - time.perf_counter())))
            elif pending:
                await asyncio.wait(pending, timeout=0.05, return_when=asyncio.FIRST_COMPLETED)
            else:
                await asyncio.sleep(0.001)
        return n_jobs
