"""CPU: pins tests/sched_standin.py -- the restatement of the reference server's dispatch rules that shards requests across
the in-process workers -- against the rules and the quirks of the reference it follows:
JobScheduler.ts:137-217 (queue pass, priority sort), :317-360 (least-loaded, then tier), WorkerRegistry.ts:261-270
(heartbeat), :325-332 (status update: a reported 0 never clears the count), :397-403 (availability), :421-462 (job counts)."""
import asyncio
import json

from gridllm_b200.worker import LocalBus
from sched_standin import SchedulerStandIn, _js_or


def _run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


def _reg(wid, models=("m",), tier=None):
    caps = {"workerId": wid, "availableModels": [{"name": n} for n in models], "maxConcurrentTasks": 1, "supportedFormats": ["json", "text"]}
    if tier:
        caps["performanceTier"] = tier
    return json.dumps({"workerId": wid, "capabilities": caps, "status": "online", "registeredAt": "t"})


def _job(i, prio="medium", model="m"):
    return {"id": f"j{i}", "model": model, "priority": prio, "timeout": 1000}


def test_javascript_or():
    assert _js_or(0, 7) == 7 and _js_or(None, 7) == 7 and _js_or("", 7) == 7
    assert _js_or(3, 7) == 3 and _js_or([], 7) == []          # an empty array is truthy: the reference heartbeat's currentJobs: []


def test_least_loaded_then_tier_and_priority_order():
    async def go():
        bus = LocalBus()
        s = SchedulerStandIn(bus, max_jobs_per_worker=2)
        await s.start()
        sent = []
        for wid, tier in (("w-low", "low"), ("w-none", None), ("w-high", "high")):
            await bus.publish("worker:registered", _reg(wid, tier=tier))

            async def on_job(msg, wid=wid):
                sent.append((wid, json.loads(msg)["job"]["jobId"]))
            await bus.subscribe(f"worker:{wid}:job", on_job)
        for i, prio in enumerate(["low", "medium", "high", "medium", None, "high", "low"]):
            j = _job(i, prio)
            if prio is None:
                del j["priority"]                                   # missing priority counts as medium
            s.add_job(j)
        await asyncio.gather(*(await s.tick()))
        return s, sent
    s, sent = _run(go())
    # queue order: high (j2, j5), medium incl. missing (j1, j3, j4), low (j0, j6) -- stable inside a class
    assert [j for _w, j in sent] == ["j2", "j5", "j1", "j3", "j4", "j0"]
    # equal load: tier decides (high, then missing = medium, then low); then the least loaded again, round and round
    assert [w for w, _j in sent] == ["w-high", "w-none", "w-low", "w-high", "w-none", "w-low"]
    assert [j["id"] for j in s.queue] == ["j6"]                # every worker holds 2 = the maximum: the seventh job waits
    assert all(w["status"] == "busy" and w["currentJobs"] == 2 for w in s.workers.values())


def test_completion_frees_a_worker_and_status_update_quirk():
    async def go():
        bus = LocalBus()
        s = SchedulerStandIn(bus, max_jobs_per_worker=1)
        await s.start()
        await bus.publish("worker:registered", _reg("w0"))
        await bus.publish("worker:registered", _reg("w1", models=("other",)))
        s.add_job(_job(0))
        s.add_job(_job(1))
        await asyncio.gather(*(await s.tick()))
        assert s.assigned == {"j0": "w0"} and [j["id"] for j in s.queue] == ["j1"]          # w1 does not list the model
        w0 = s.workers["w0"]
        assert w0["currentJobs"] == 1 and w0["status"] == "busy"
        # the worker reports currentJobs 0 in a status update: `data.currentJobs || worker.currentJobs` keeps the 1
        await bus.publish("worker:status_update", json.dumps({"workerId": "w0", "status": "online", "currentJobs": 0}))
        assert w0["currentJobs"] == 1 and w0["status"] == "online"
        assert s.select(_job(9)) is None                            # online, but still holding its one job
        # a heartbeat, on the other hand, takes 0 as 0 (`data.currentJobs || 0`) -- and the reference client's `[]` verbatim
        await bus.publish("worker:heartbeat", json.dumps({"workerId": "w0", "status": "online", "currentJobs": [], "timestamp": "t"}))
        assert w0["currentJobs"] == [] and s.select(_job(9))["workerId"] == "w0"           # [] < 1 in JS: available again
        await bus.publish("worker:heartbeat", json.dumps({"workerId": "w0", "status": "busy", "currentJobs": 1, "timestamp": "t"}))
        assert s.select(_job(9)) is None
        # only job completion frees the worker in the server's own bookkeeping
        await bus.publish("job:completed", json.dumps({"jobId": "j0", "workerId": "w0", "result": {}, "timestamp": "t"}))
        assert w0["currentJobs"] == 0 and w0["status"] == "online" and "j0" in s.results
        await asyncio.gather(*(await s.tick()))
        assert s.assigned["j1"] == "w0"
        return s
    _run(go())


def test_cancellation_message_reaches_the_holder():
    async def go():
        bus = LocalBus()
        s = SchedulerStandIn(bus)
        await s.start()
        await bus.publish("worker:registered", _reg("w0"))
        got = []

        async def on_job(msg):
            got.append(json.loads(msg))
        await bus.subscribe("worker:w0:job", on_job)
        s.add_job(_job(0))
        await asyncio.gather(*(await s.tick()))
        await s.cancel_job("j0")
        return got
    got = _run(go())
    assert got[0]["type"] == "job_assignment" and got[0]["job"]["jobId"] == "j0" and got[0]["job"]["timeout"] == 1000
    assert got[1] == {"type": "job_cancellation", "jobId": "j0"}
