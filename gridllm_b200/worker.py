"""Worker host around the native engine: the job-dispatch half of the reference's
``WorkerClientService`` (/root/reference/client/src/services/WorkerClientService.ts), speaking the
frozen pub/sub wire format to whatever bus the server side uses.

What is mirrored (and what is deliberately not):
  * registration object + ``worker:registered``                     WorkerClientService.ts:175-206
  * heartbeat object + ``worker:heartbeat`` / ``heartbeat:<id>``      :330-353   (currentJobs sent as a number,
                                                                    accepted by WorkerRegistry.ts:269)
  * ``worker:status_update``                                         :442-472
  * ``worker:<id>:job`` {type: job_assignment} dispatch by request type, stream chunk re-publish on
    ``job:stream:<jobId>``, ``job:completed`` / ``job:failed`` / ``job:result:<jobId>``   :497-712
  * busy-drop of a second assignment                                 :500-505
  * ``job_cancellation`` IS honoured here (the reference ignores it, :483-491)
The Redis connection manager, Express health server, Joi config and winston logging of the reference
client are out of scope: the bus is an interface with an in-process
implementation; a Redis-backed one plugs in the same way on a box that has Redis.
"""
from __future__ import annotations

import asyncio
import json
from datetime import datetime, timezone
from typing import Any, Awaitable, Callable, Dict, List, Optional

from .service import NativeInferenceService


def _iso() -> str:
    return datetime.now(timezone.utc).isoformat(timespec="milliseconds").replace("+00:00", "Z")


class LocalBus:
    """In-process stand-in for the Redis pub/sub + hash/string keys the reference uses
    (server/src/services/RedisService.ts:111-227).  Channel names travel bare, keys get no prefix."""

    def __init__(self):
        self.subs: Dict[str, List[Callable[[str], Awaitable[None]]]] = {}
        self.hashes: Dict[str, Dict[str, str]] = {}
        self.keys: Dict[str, str] = {}
        self.log: List[tuple] = []           # (channel, message) in publish order -- golden-fixture tests read this

    async def publish(self, channel: str, message: str) -> None:
        self.log.append((channel, message))
        for cb in list(self.subs.get(channel, [])):
            await cb(message)

    async def subscribe(self, channel: str, cb: Callable[[str], Awaitable[None]]) -> None:
        self.subs.setdefault(channel, []).append(cb)

    async def hset(self, key: str, field: str, value: str) -> None:
        self.hashes.setdefault(key, {})[field] = value

    async def set_with_expiry(self, key: str, value: str, ttl_s: float) -> None:
        self.keys[key] = value


class NativeWorker:
    """One worker id <-> one GPU engine."""

    def __init__(self, worker_id: str, service: NativeInferenceService, bus: LocalBus, heartbeat_interval_ms: int = 5000,
                 max_concurrent: int = 1):
        self.worker_id = worker_id
        self.service = service
        self.bus = bus
        self.heartbeat_interval_ms = heartbeat_interval_ms
        # jobs this worker holds at once (MAX_CONCURRENT_JOBS_PER_WORKER on the server side, server/src/config/index.ts:31).
        # 1 = the reference's busy-drop (:500-505).  More than one: the jobs run as concurrent tasks, and a service created with
        # max_batch > 1 decodes them TOGETHER (continuous batching, gridllm_b200/batching.py) -- one batched step per token for
        # every job the worker holds; with an un-batched service they wait at the engine one after the other
        self.max_concurrent = max(1, int(max_concurrent))
        self._tasks: set = set()
        self.isProcessingJob = False
        self.currentJobs = 0
        self.capabilities: Optional[Dict[str, Any]] = None
        self._cancelled: set = set()
        self._hb_task: Optional[asyncio.Task] = None

    # ---- boot (WorkerClientService.initialize/start, :35-89) ----------------------------------------
    async def start(self) -> None:
        if not await self.service.checkHealth():
            raise RuntimeError("Ollama service is not available")      # same message the reference throws (:43-47)
        models = await self.service.getAvailableModels()
        self.capabilities = {"workerId": self.worker_id, "availableModels": models, "maxConcurrentTasks": self.max_concurrent,
                             "supportedFormats": ["json", "text"], "lastUpdated": _iso()}
        reg = {"workerId": self.worker_id, "capabilities": self.capabilities, "status": "online", "registeredAt": _iso()}
        await self.bus.hset("workers", self.worker_id, json.dumps(reg))
        await self.bus.publish("worker:registered", json.dumps(reg))
        await self.bus.subscribe(f"worker:{self.worker_id}:job", self.handleJobMessage)

    def _status(self) -> str:
        """The server takes `status` verbatim (WorkerRegistry.ts:266,331) and only assigns to workers that are "online"
        (WorkerRegistry.ts:397-403).  Its own bookkeeping calls a worker "busy" once currentJobs reaches
        MAX_CONCURRENT_JOBS_PER_WORKER (WorkerRegistry.ts:431-438) -- so a worker that can hold several jobs reports "busy" only
        when it is full; with max_concurrent 1 this is the reference's busy-while-processing (WorkerClientService.ts:337,451)."""
        return "busy" if self.currentJobs >= self.max_concurrent else "online"

    async def sendHeartbeat(self) -> None:
        hb = {"workerId": self.worker_id, "status": self._status(), "timestamp": _iso(),
              "currentJobs": self.currentJobs, "connectionHealth": "healthy"}
        await self.bus.set_with_expiry(f"heartbeat:{self.worker_id}", json.dumps(hb), self.heartbeat_interval_ms * 2 / 1000)
        await self.bus.publish("worker:heartbeat", json.dumps(hb))

    # the reference fires heartbeats from a timer (WorkerClientService.ts:316-323); they must keep firing while a job runs,
    # which is why every engine call of the service runs off the event loop (asyncio.to_thread / a worker thread)
    def start_heartbeats(self) -> None:
        async def loop():
            while True:
                await self.sendHeartbeat()
                await asyncio.sleep(self.heartbeat_interval_ms / 1000)
        if self._hb_task is None:
            self._hb_task = asyncio.get_running_loop().create_task(loop())

    async def drain(self) -> None:
        """wait for the jobs this worker holds (max_concurrent > 1 runs them as tasks)"""
        # A task that has finished leaves `_tasks` through its done-callback, which runs one loop iteration AFTER the task completes.
        # gather() of tasks that are all done already returns a completed future and awaiting it does not yield: without the explicit
        # pruning and the sleep(0) below this loop spun for ever when drain() was entered in that one-iteration window (found on two
        # GPUs with the 1 s dispatch tick: the event loop never got to run the callbacks).
        while self._tasks:
            await asyncio.gather(*list(self._tasks), return_exceptions=True)
            self._tasks = {t for t in self._tasks if not t.done()}
            await asyncio.sleep(0)

    async def stop(self) -> None:
        await self.drain()
        if self._hb_task is not None:
            self._hb_task.cancel()
            try:
                await self._hb_task
            except asyncio.CancelledError:
                pass
            self._hb_task = None

    async def publishStatusUpdate(self) -> None:
        await self.bus.publish("worker:status_update", json.dumps(
            {"workerId": self.worker_id, "status": self._status(), "currentJobs": self.currentJobs}))

    # ---- message handling (:478-495) -----------------------------------------------------------------
    async def handleJobMessage(self, message: str) -> None:
        data = json.loads(message)
        if data.get("type") == "job_assignment":
            if self.max_concurrent == 1:
                await self.processJobAssignment(data["job"])
            else:                                    # run beside the jobs already held; the message handler returns at once
                # the job is COUNTED here, synchronously with the message: the count the next status update / heartbeat reports
                # (which the server takes over verbatim, WorkerRegistry.ts:269,332) never lags behind what was accepted
                if self.currentJobs >= self.max_concurrent:
                    return                           # full: dropped, like the reference's busy-drop (:500-505)
                self.currentJobs += 1
                self.isProcessingJob = True
                t = asyncio.get_running_loop().create_task(self.processJobAssignment(data["job"], counted=True))
                self._tasks.add(t)
                t.add_done_callback(self._tasks.discard)
        elif data.get("type") == "job_cancellation":
            self._cancelled.add(data.get("jobId"))

    # ---- processJobAssignment (:497-712) ---------------------------------------------------------------
    async def processJobAssignment(self, assignment: Dict[str, Any], counted: bool = False) -> None:
        request = assignment["request"]
        if not counted:
            if self.currentJobs >= self.max_concurrent or (self.max_concurrent == 1 and self.isProcessingJob):
                return                               # dropped; the server notices via timeout / orphan scan
            self.currentJobs += 1
            self.isProcessingJob = True
        await self.publishStatusUpdate()
        jid = request["id"]
        try:
            if not await self.service.validateModel(request["model"]):
                raise RuntimeError(f"Model {request['model']} is not available")
            md = request.get("metadata") or {}
            rtype = md.get("requestType")
            result: Optional[Dict[str, Any]] = None
            if rtype == "embedding":
                result = await self.service.generateEmbedding(request)
            elif rtype == "chat":
                if request.get("stream"):
                    full = ""
                    agen = self.service.generateChatStreamResponse(request)
                    try:
                        async for chunk in agen:
                            full += chunk["response"]
                            await self.bus.publish(f"job:stream:{jid}", json.dumps(
                                {"jobId": jid, "workerId": self.worker_id,
                                 "chunk": dict(chunk, message={"content": chunk["response"]}), "timestamp": _iso()}))
                            if chunk["done"] or jid in self._cancelled:
                                result = dict(chunk, id=jid, message={"content": full})
                                break
                    finally:
                        await agen.aclose()          # a cancelled job: closing the generator stops the engine at its next token
                else:
                    result = await self.service.generateChatResponse(request)
            elif request.get("stream"):
                full = ""
                agen = self.service.generateStreamResponse(request)
                try:
                    async for chunk in agen:
                        full += chunk["response"]
                        await self.bus.publish(f"job:stream:{jid}", json.dumps(
                            {"jobId": jid, "workerId": self.worker_id,
                             "chunk": {"id": chunk["id"], "response": chunk["response"], "done": chunk["done"]}, "timestamp": _iso()}))
                        if chunk["done"] or jid in self._cancelled:
                            result = {"id": jid, "response": full, "done": True}
                            break
                finally:
                    await agen.aclose()
            else:
                result = await self.service.generateResponse(request)
            if not result:
                raise RuntimeError("No result generated from inference")
            payload = json.dumps({"jobId": jid, "workerId": self.worker_id, "result": result, "timestamp": _iso()})
            await self.bus.publish("job:completed", payload)
            await self.bus.publish(f"job:result:{jid}", payload)
        except Exception as error:
            msg = str(error) or "Unknown error"
            payload = json.dumps({"jobId": jid, "workerId": self.worker_id, "error": msg, "timestamp": _iso()})
            await self.bus.publish("job:failed", payload)
            await self.bus.publish(f"job:result:{jid}", payload)
        finally:
            self.currentJobs -= 1
            self.isProcessingJob = self.currentJobs > 0
            self._cancelled.discard(jid)
            await self.publishStatusUpdate()
