"""TEST INFRASTRUCTURE: a stand-in for the reference server's dispatch loop, restating only the selection
rule the north_star says is 'reused unchanged' -- JobScheduler.selectWorkerForJob
(/root/reference/server/src/services/JobScheduler.ts:317-360: available workers that list the model,
fewest currentJobs first, then performance tier) over WorkerRegistry availability
(server/src/services/WorkerRegistry.ts:397-403: status online and currentJobs < max, default 1) and the
priority sort of processJobQueue (JobScheduler.ts:145-151).  The real server is TypeScript + Redis and cannot
run in this image; this drives N in-process NativeWorkers through the same pub/sub messages."""
import asyncio
import json
from datetime import datetime, timezone

PRIO = {"high": 3, "medium": 2, "low": 1}


class SchedulerStandIn:
    def __init__(self, bus, max_jobs_per_worker=1):
        self.bus = bus
        self.workers = {}
        self.queue = []
        self.results = {}
        self.assigned = {}
        self.max_jobs = max_jobs_per_worker

    async def start(self):
        await self.bus.subscribe("worker:registered", self._on_registered)
        await self.bus.subscribe("worker:status_update", self._on_status)
        await self.bus.subscribe("job:completed", self._on_done)
        await self.bus.subscribe("job:failed", self._on_done)

    async def _on_registered(self, msg):
        d = json.loads(msg)
        self.workers[d["workerId"]] = {"workerId": d["workerId"], "capabilities": d["capabilities"], "status": "online", "currentJobs": 0}

    async def _on_status(self, msg):
        d = json.loads(msg)
        w = self.workers.get(d["workerId"])
        if w:
            w["status"] = d["status"]
            w["currentJobs"] = d["currentJobs"] or w["currentJobs"]     # WorkerRegistry.ts:332 quirk kept

    async def _on_done(self, msg):
        d = json.loads(msg)
        self.results[d["jobId"]] = d
        w = self.workers.get(d["workerId"])
        if w:
            w["currentJobs"] = 0                                        # markWorkerAvailable, WorkerRegistry.ts:460-462
            w["status"] = "online"

    def select(self, job):
        cand = [w for w in self.workers.values() if w["status"] == "online" and w["currentJobs"] < self.max_jobs
                and any(m["name"] == job["model"] for m in w["capabilities"]["availableModels"])]
        cand.sort(key=lambda w: (w["currentJobs"], -2))
        return cand[0] if cand else None

    def add_job(self, request):
        self.queue.append(request)

    async def tick(self):
        """one pass of processJobQueue: priority order, assign while workers are available"""
        self.queue.sort(key=lambda j: -PRIO.get(j.get("priority", "medium"), 2))
        launched = []
        rest = []
        for job in self.queue:
            w = self.select(job)
            if w is None:
                rest.append(job)
                continue
            w["currentJobs"] += 1                                       # markWorkerBusy
            self.assigned[job["id"]] = w["workerId"]
            msg = json.dumps({"type": "job_assignment", "job": {"jobId": job["id"], "workerId": w["workerId"], "request": job,
                                                                "assignedAt": datetime.now(timezone.utc).isoformat(), "timeout": 300000}})
            launched.append(asyncio.ensure_future(self.bus.publish(f"worker:{w['workerId']}:job", msg)))
        self.queue = rest
        return launched

    async def run_until_empty(self):
        pending = []
        while self.queue or pending:
            pending += await self.tick()
            if pending:
                done, not_done = await asyncio.wait(pending, return_when=asyncio.FIRST_COMPLETED)
                pending = list(not_done)
            else:
                await asyncio.sleep(0)
