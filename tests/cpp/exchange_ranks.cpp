// exchange_ranks.cpp — the multi-GPU host layer driven from C++ only: N processes (fork), a socket
// hub as the host all-gather, Zenyth::PeerExchange over real CUDA IPC.  No torch, no NCCL, no Python.
//
//     exchange_ranks [world=2] [scenes_per_rank=1] [frames=9] [levels=7] [overlap=0] [worklist=0]
//
// Rank r runs on GPU r % device_count, so with ONE GPU both ranks share cuda:0 (CUDA IPC between
// processes on the same device is legal): IPC mapping, peer stores, flag stores and acquire loads
// and the 4-slot / 2-list rotation are exercised exactly as across NVLink.  Ranks that SHARE a GPU
// run in lockstep (stream sync + host rendezvous after every frame), so that no kernel ever has to
// spin on a flag another process on the same GPU has yet to raise — kernels of different processes
// on one device are not guaranteed to run at the same time, and a spinning kernel can starve the
// one it waits for (Xid 109 on this driver).  With one GPU per rank nothing is synchronised on the
// host and the flags are waited on by the device, as in production.  Every frame uses another camera,
// and every rank checks EVERY gathered list of EVERY frame against the list a plain single-GPU
// update of that scene produces.  The parent never touches CUDA (fork before any CUDA call).
#include "zenyth_b200.hpp"

#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

using namespace Zenyth;

static bool read_all(int fd, void* p, size_t n) {
	auto* c = static_cast<char*>(p);
	while (n) {
		const ssize_t k = read(fd, c, n);
		if (k <= 0) return false;
		c += k; n -= size_t(k);
	}
	return true;
}
static bool write_all(int fd, const void* p, size_t n) {
	auto* c = static_cast<const char*>(p);
	while (n) {
		const ssize_t k = write(fd, c, n);
		if (k <= 0) return false;
		c += k; n -= size_t(k);
	}
	return true;
}

// hub (parent): one round = a u64 size + payload from every rank, the concatenation back to every rank
static void hub(const std::vector<int>& fds) {
	for (;;) {
		std::vector<char> all;
		uint64_t bytes = 0;
		for (size_t r = 0; r < fds.size(); ++r) {
			uint64_t n = 0;
			if (!read_all(fds[r], &n, 8)) return;      // a rank closed its socket: done
			if (r == 0) { bytes = n; all.resize(size_t(n) * fds.size()); }
			if (n != bytes || !read_all(fds[r], all.data() + r * size_t(n), size_t(n))) return;
		}
		for (int fd : fds)
			if (!write_all(fd, all.data(), all.size())) return;
	}
}

struct Channel { int fd; uint32_t world; };
static int all_gather(void* user, const void* send, void* recv, size_t bytes) {
	auto* ch = static_cast<Channel*>(user);
	const uint64_t n = bytes;
	if (!write_all(ch->fd, &n, 8) || !write_all(ch->fd, send, bytes)) return 1;
	return read_all(ch->fd, recv, bytes * ch->world) ? 0 : 1;
}

static CameraDesc camera_of_frame(int k) {
	CameraDesc c;
	c.eye = {20.0f * float(k), -10.0f * float(k), 1500.0f - 100.0f * float(k)};
	c.center = {c.eye.x, c.eye.y, c.eye.z - 1.0f};
	return c;
}

static int rank_main(uint32_t rank, uint32_t world, uint32_t S, int frames, int levels, bool overlap, bool worklist, int fd) {
	try {
		int devices = 0;
		if (zn_device_count(&devices) != ZN_OK || devices < 1) { std::fprintf(stderr, "rank %u: no CUDA device\n", rank); return 3; }
		GpuContext ctx(int(rank) % devices);
		const bool lockstep = uint32_t(devices) < world;     // some ranks share a GPU
		SceneDesc desc;
		desc.levelOffsets.push_back(0);
		for (int l = 0, n = 7; l < levels; ++l, n *= 8) desc.levelOffsets.push_back(desc.levelOffsets.back() + uint32_t(n));
		const uint32_t E = desc.levelOffsets.back();
		desc.entityCount = E;
		const uint32_t total = world * S;
		std::vector<uint32_t> bases = SceneIndexBases(std::vector<uint32_t>(total, E));
		const std::vector<uint32_t> mine = ShardScenes(total, world, rank);
		if (mine.size() != S || mine[0] != rank * S) { std::fprintf(stderr, "rank %u: ShardScenes\n", rank); return 4; }

		// expected lists: every scene of every rank, every frame's camera, by the plain single-GPU path
		std::vector<std::vector<std::vector<uint32_t>>> expect;
		expect.assign(static_cast<size_t>(frames), std::vector<std::vector<uint32_t>>(total));
		{
			Scene plain(ctx, desc);
			for (uint32_t g = 0; g < total; ++g) {
				detail::Check(zn_scene_fill_synthetic(plain.Handle(), 0x5A454E59u + 11u + g, 8, 0.25f), "fill");
				plain.SetIndexBase(bases[g]);
				plain.UpdateWorld();
				for (int k = 0; k < frames; ++k) {
					plain.Cull(Camera(camera_of_frame(k)));
					expect[size_t(k)][g] = plain.Visible();
				}
			}
		}
		std::vector<std::unique_ptr<Scene>> scenes;
		std::vector<Scene*> ptrs;
		for (uint32_t j = 0; j < S; ++j) {
			scenes.push_back(std::make_unique<Scene>(ctx, desc));
			detail::Check(zn_scene_fill_synthetic(scenes.back()->Handle(), 0x5A454E59u + 11u + mine[j], 8, 0.25f), "fill");
			scenes.back()->SetIndexBase(bases[mine[j]]);
			scenes.back()->SetExpandWorkList(worklist);   // the expansion strategy of the scene the exchange expands with
			ptrs.push_back(scenes.back().get());
		}
		Channel ch{fd, world};
		PeerExchangeDesc xd;
		xd.worldSize = world; xd.rank = rank; xd.recordsPerRank = S; xd.allGather = all_gather; xd.user = &ch; xd.indexBases = bases;
		xd.timeoutMs = 5000;
		xd.overlap = overlap;
		{
			PeerExchange exchange(ctx, *scenes[0], xd);
			std::vector<uint32_t> host;
			auto verify = [&](int k) {
				const PeerExchange::Lists got = exchange.Result();
				for (uint32_t g = 0; g < total; ++g) {
					const std::vector<uint32_t>& want = expect[size_t(k)][g];
					if (got.counts[g] != want.size()) {
						std::fprintf(stderr, "rank %u frame %d scene %u: count %u, expected %zu\n", rank, k, g, got.counts[g], want.size());
						return false;
					}
					host.resize(want.size());
					ctx.Read(got.device + size_t(g) * got.stride, host.data(), want.size() * 4);
					if (std::memcmp(host.data(), want.data(), want.size() * 4) != 0) {
						std::fprintf(stderr, "rank %u frame %d scene %u: list differs from the single-GPU result\n", rank, k, g);
						return false;
					}
				}
				return true;
			};
			auto rendezvous = [&]() {
				uint32_t token = rank;
				std::vector<uint32_t> all(world);
				return all_gather(&ch, &token, all.data(), 4) == 0;
			};
			// default: Expand one frame behind, on the context stream; overlap: Expand the frame just updated, on the side stream
			const int lag = overlap ? 0 : 1;
			for (int k = 0; k < frames; ++k) {
				const Camera cam(camera_of_frame(k));
				exchange.Bind(ptrs, uint64_t(k));
				for (Scene* s : ptrs) { s->SetCamera(cam); s->Update(); }
				if (lockstep) {
					// shared GPU: raise the flags (no waiting), make sure every rank has, and only then launch a kernel that waits
					exchange.Signal();
					if (!rendezvous()) return 6;
				}
				exchange.Expand(k - lag);
				if (k - lag >= 0 && (lockstep || k % 2 == 1)) {   // every frame in lockstep; every other one otherwise, so that frames really pipeline
					if (!verify(k - lag)) return 5;
				}
			}
			exchange.Expand(frames - 1);
			if (!verify(frames - 1)) return 5;
			size_t visible = 0;
			for (const auto& v : expect[size_t(frames - 1)]) visible += v.size();
			std::printf("exchange ok rank %u/%u: %u scenes x %u entities, %d frames, last frame %zu visible\n", rank, world, total, E, frames, visible);
		}   // ~PeerExchange: collective teardown
		close(fd);
		return 0;
	} catch (const std::exception& e) {
		std::fprintf(stderr, "rank %u: %s\n", rank, e.what());
		return 2;
	}
}

int main(int argc, char** argv) {
	const uint32_t world = argc > 1 ? uint32_t(std::atoi(argv[1])) : 2u;
	const uint32_t S = argc > 2 ? uint32_t(std::atoi(argv[2])) : 1u;
	const int frames = argc > 3 ? std::atoi(argv[3]) : 9;
	const int levels = argc > 4 ? std::atoi(argv[4]) : 7;
	const bool overlap = argc > 5 && std::atoi(argv[5]) != 0;
	const bool worklist = argc > 6 && std::atoi(argv[6]) != 0;
	std::vector<int> hub_fds;
	std::vector<pid_t> kids;
	for (uint32_t r = 0; r < world; ++r) {
		int sv[2];
		if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { std::perror("socketpair"); return 1; }
		const pid_t pid = fork();
		if (pid == 0) {
			for (int fd : hub_fds) close(fd);
			close(sv[0]);
			std::exit(rank_main(r, world, S, frames, levels, overlap, worklist, sv[1]));
		}
		close(sv[1]);
		hub_fds.push_back(sv[0]);
		kids.push_back(pid);
	}
	hub(hub_fds);
	int bad = 0;
	for (pid_t pid : kids) {
		int st = 0;
		waitpid(pid, &st, 0);
		if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++bad;
	}
	for (int fd : hub_fds) close(fd);
	return bad ? 1 : 0;
}
