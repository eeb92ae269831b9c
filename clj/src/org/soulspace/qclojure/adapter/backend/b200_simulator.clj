(ns org.soulspace.qclojure.adapter.backend.b200-simulator
  "B200 state-vector backend for QClojure: a QuantumBackend (application/backend.clj:72-112) whose execution path is
  libqcb200.so (include/qcb200.h) instead of domain/circuit/execute-circuit.

  NOT COMPILED OR RUN IN THIS REPOSITORY'S CI: the build image has no JVM (see DESIGN.md §1).  The record mirrors
  LocalQuantumSimulator (adapter/backend/ideal_simulator.clj:100-176) method by method; the same logic is exercised
  through the Python mirror qclojure_b200/backend.py, which calls the identical C entry points.

  Binding: Java FFM (JDK 22+, java.lang.foreign).  On older JVMs replace `ffi` with the JNA variant at the bottom.
  Every C function returns int32 status; 0 = ok, message via qcb_last_error."
  (:require [org.soulspace.qclojure.application.backend :as backend]
            [org.soulspace.qclojure.domain.operation-registry :as opreg]
            [fastmath.complex :as fc])
  (:import (java.lang.foreign Arena FunctionDescriptor Linker MemoryLayout MemorySegment SymbolLookup ValueLayout)
           (java.lang.invoke MethodHandle)))

;;; ------------------------------------------------------------------ FFM plumbing
(def ^:private ^Linker linker (Linker/nativeLinker))
(def ^:private lib-arena (Arena/global))
(defonce ^:private ^SymbolLookup lookup
  (delay (SymbolLookup/libraryLookup (or (System/getProperty "qcb200.lib") "libqcb200.so") lib-arena)))

(def ^:private I32 ValueLayout/JAVA_INT)
(def ^:private I64 ValueLayout/JAVA_LONG)
(def ^:private F64 ValueLayout/JAVA_DOUBLE)
(def ^:private PTR ValueLayout/ADDRESS)

(defn- ^MethodHandle ffi
  "Downcall handle for `int32_t name(args...)`."
  [name & arg-layouts]
  (.downcallHandle linker (.orElseThrow (.find ^SymbolLookup @lookup name))
                   (FunctionDescriptor/of I32 (into-array MemoryLayout arg-layouts))))

;; include/qcb200.h — one handle per entry point the hot path needs
(def ^:private qcb-create        (delay (ffi "qcb_create" PTR PTR)))                 ; (const qcb_config*, qcb_handle*)
(def ^:private qcb-config-default (delay (ffi "qcb_config_default" PTR)))
(def ^:private qcb-destroy       (delay (ffi "qcb_destroy" PTR)))
(def ^:private qcb-last-error    (delay (ffi "qcb_last_error" PTR PTR I64)))
(def ^:private qcb-submit        (delay (ffi "qcb_submit" PTR PTR PTR)))             ; (h, const qcb_job_request*, uint64* id)
(def ^:private qcb-job-status    (delay (ffi "qcb_job_status" PTR I64 PTR)))
(def ^:private qcb-job-result    (delay (ffi "qcb_job_result_get" PTR I64 PTR)))
(def ^:private qcb-cancel        (delay (ffi "qcb_cancel" PTR I64 PTR)))
(def ^:private qcb-queue-status  (delay (ffi "qcb_queue_status" PTR PTR PTR PTR)))

;;; ------------------------------------------------------------------ circuit map -> qcb_op[]
;; struct qcb_op { int32 kind; int32 q[3]; int32 n_mask; int32 _pad; uint64 mask; double angle; double mat[8]; void* ext; } = 112 bytes
(def ^:private op-size 112)

(def ^:private kind-code
  "enum qcb_op_kind (include/qcb200.h) for every branch of apply-gate-to-state (domain/circuit.clj:964-1071)."
  (zipmap [:i :x :y :z :h :s :s-dag :t :t-dag :rx :ry :rz :phase :cnot :cz :cy :crx :cry :crz :swap :iswap :toffoli
           :fredkin :rydberg-cz :rydberg-cphase :rydberg-blockade :global-h :global-x :global-y :global-z
           :global-rx :global-ry :global-rz]
          (range)))

(defn- operands
  "The operand order the C ABI expects, from the reference's :operation-params keys (circuit.clj:170-845).
  A missing :target defaults to qubit 0 exactly like circuit.clj:965-984."
  [op-type p]
  (case op-type
    (:cnot :cz :cy :crx :cry :crz :rydberg-cz :rydberg-cphase) [(:control p) (:target p) -1]
    (:swap :iswap) [(:qubit1 p) (:qubit2 p) -1]
    :toffoli [(:control1 p) (:control2 p) (:target p)]
    :fredkin [(:control p) (:target1 p) (:target2 p)]
    [(get p :target 0) -1 -1]))

(defn- write-op! [^MemorySegment seg idx {:keys [operation-type operation-params]}]
  (let [t (opreg/resolve-gate-alias operation-type)           ; aliases resolved like circuit.clj:953
        base (* idx op-size)
        [q0 q1 q2] (operands t operation-params)
        qs (:qubit-indices operation-params)]
    (when-not (kind-code t) (throw (ex-info "Unknown gate type" {:operation-type operation-type})))
    (.set seg I32 (long base) (int (kind-code t)))
    (.set seg I32 (long (+ base 4)) (int q0)) (.set seg I32 (long (+ base 8)) (int q1)) (.set seg I32 (long (+ base 12)) (int q2))
    (.set seg I32 (long (+ base 16)) (int (count qs)))
    (.set seg I64 (long (+ base 24)) (long (reduce (fn [m q] (bit-or m (bit-shift-left 1 q))) 0 qs)))
    (.set seg F64 (long (+ base 32)) (double (get operation-params :angle 0.0)))))

(defn- encode-ops ^MemorySegment [^Arena arena operations]
  (let [gates (remove #(= :measure (:operation-type %)) operations)   ; final measurement = shots, like the reference
        seg (.allocate arena (long (* op-size (max 1 (count gates)))) 8)]
    (doseq [[i op] (map-indexed vector gates)] (write-op! seg i op))
    [seg (count gates)]))

;;; ------------------------------------------------------------------ the backend record
(defonce ^:private job-table (atom {}))          ; job-id string -> {:handle :native-id :n :shots :specs}

(defn- check! [handle rc]
  (when-not (zero? rc)
    (with-open [a (Arena/ofConfined)]
      (let [buf (.allocate a 512)]
        (.invokeWithArguments ^MethodHandle @qcb-last-error [handle buf (long 512)])
        (throw (ex-info (.getString buf 0) {:qcb-status rc}))))))

(defn- open-handle
  "qcb_create for n qubits (one handle = one state vector resident in HBM; cached per qubit count by the record)."
  [n {:keys [device strict-parity] :or {device -1 strict-parity 1}}]
  (with-open [a (Arena/ofConfined)]
    (let [cfg (.allocate a 96 8) out (.allocate a 8 8)]
      (.invokeWithArguments ^MethodHandle @qcb-config-default [cfg])
      (.set cfg I32 0 (int n)) (.set cfg I32 4 (int device)) (.set cfg I32 12 (int strict-parity))
      (check! MemorySegment/NULL (.invokeWithArguments ^MethodHandle @qcb-create [cfg out]))
      (.get out PTR 0))))

(defrecord B200Simulator [config handles]
  backend/QuantumBackend
  (backend-info [_]
    {:backend-type :simulator
     :backend-name "B200 state-vector simulator (libqcb200)"
     :description "fp64 state-vector simulation on NVIDIA B200: fused shared-memory gate sweeps, fp64 tensor-core rounds"
     :backend-config config
     :max-qubits (get config :max-qubits 33)
     :capabilities #{:quantum-backend}
     :device (:device config)
     :version "0.1.0"})
  (device [_] (:device config))
  (available? [_] true)

  (submit-circuit [_ circuit options]
    ;; same contract as ideal_simulator.clj:119-136: returns a job-id string immediately; the native worker thread runs the job
    (let [n (:num-qubits circuit)
          handle (or (get @handles n) (get (swap! handles #(if (% n) % (assoc % n (open-handle n config)))) n))
          specs (:result-specs options)
          shots (get-in specs [:measurements :shots] 0)
          uniforms (double-array (repeatedly shots rand))        ; the reference draws Math/random per shot (state.clj:903)
          job-id (str "b200_job_" (System/nanoTime))]
      (with-open [a (Arena/ofConfined)]
        (let [[ops n-ops] (encode-ops a (:operations circuit))
              useg (.allocateFrom a F64 uniforms)
              req (.allocate a 88 8) idseg (.allocate a 8 8)]
          ;; struct qcb_job_request (include/qcb200.h): ops, n_ops, initial_state, initial_count, uniforms, n_shots, ham..., flags
          (.set req PTR 0 ops) (.set req I64 8 (long n-ops))
          ;; :initial-state (ideal_simulator.clj:84-86): interleaved re/im doubles, copied by qcb_submit
          (if-let [init (:initial-state options)]
            (let [amps (:state-vector init)
                  iseg (.allocate a (long (* 16 (count amps))) 8)]
              (doseq [[i z] (map-indexed vector amps)]
                (.setAtIndex iseg F64 (long (* 2 i)) (double (fc/re z)))
                (.setAtIndex iseg F64 (long (inc (* 2 i))) (double (fc/im z))))
              (.set req PTR 16 iseg) (.set req I64 24 (long (count amps))))
            (do (.set req PTR 16 MemorySegment/NULL) (.set req I64 24 0)))
          (.set req PTR 32 useg) (.set req I64 40 (long shots))
          ;; :hamiltonian spec = collection of {:coefficient c :pauli-string "XIZ..."} (domain/hamiltonian.clj:35-62), the
          ;; form the variational objective sends (variational_algorithm.clj:345); the noisy path's {:hamiltonian H} too
          (when-let [ham (let [h (:hamiltonian specs)] (if (map? h) (:hamiltonian h) h))]
            (let [cseg (.allocateFrom a F64 (double-array (map :coefficient ham)))
                  pseg (.allocate a (long (* 8 (count ham))) 8)]
              (doseq [[i term] (map-indexed vector ham)]
                (.setAtIndex pseg PTR (long i) (.allocateFrom a ^String (:pauli-string term))))
              (.set req PTR 48 cseg) (.set req PTR 56 pseg) (.set req I64 64 (long (count ham)))))
          (.set req I32 72 (int (if (<= n 24) 1 0)))             ; probabilities only where a Clojure vector can hold them
          (.set req I32 76 (int (if (<= n 24) 1 0)))
          (check! handle (.invokeWithArguments ^MethodHandle @qcb-submit [handle req idseg]))
          (swap! job-table assoc job-id {:handle handle :native-id (.get idseg I64 0) :n n :shots shots :specs specs})))
      job-id))

  (job-status [_ job-id]
    (if-let [{:keys [handle native-id]} (@job-table job-id)]
      (with-open [a (Arena/ofConfined)]
        (let [s (.allocate a 4 4)]
          (.invokeWithArguments ^MethodHandle @qcb-job-status [handle (long native-id) s])
          (nth [:queued :running :completed :failed :cancelled :not-found] (.get s I32 0))))
      :not-found))

  (job-result [this job-id]
    ;; result map shaped like result.clj:201-252 (ideal path): :measurement-results {:measurement-outcomes :frequencies ...}
    (if-let [{:keys [handle native-id n shots]} (@job-table job-id)]
      (if (= :completed (backend/job-status this job-id))
        (with-open [a (Arena/ofConfined)]
          (let [outc (.allocate a (long (* 8 (max 1 shots))) 8)
                res (.allocate a 344 8)]
            (.set res I64 16 (long shots)) (.set res PTR 24 outc)
            (check! handle (.invokeWithArguments ^MethodHandle @qcb-job-result [handle (long native-id) res]))
            (let [outcomes (vec (for [i (range shots)] (.getAtIndex outc I64 (long i))))
                  freq (frequencies outcomes)
                  specs (:specs (@job-table job-id))]
              {:job-id job-id :job-status :completed
               :execution-time-ms (.get res F64 8)
               :results (cond-> {:result-types (set (keys specs))}
                          (:measurements specs)
                          (assoc :measurement-results {:measurement-outcomes outcomes
                                                       :frequencies freq
                                                       :empirical-probabilities (into {} (map (fn [[k v]] [k (/ v (max 1 shots))]) freq))
                                                       :shot-count shots
                                                       :measurement-qubits (or (get-in specs [:measurements :qubits]) (range n))
                                                       :source :ideal-simulation})
                          ;; qcb_job_result.energy / has_energy (offsets 32 / 40): result.clj:327-344
                          (pos? (.get res I32 40))
                          (assoc :hamiltonian-result {:energy-expectation (.get res F64 32)
                                                      :hamiltonian (let [h (:hamiltonian specs)] (if (map? h) (:hamiltonian h) h))}))})))
        {:job-id job-id :job-status (backend/job-status this job-id) :error-message "Job not completed"})
      {:job-id job-id :job-status :not-found :error-message "Job not found"}))

  (cancel-job [_ job-id]
    (if-let [{:keys [handle native-id]} (@job-table job-id)]
      (with-open [a (Arena/ofConfined)]
        (let [s (.allocate a 4 4)]
          (.invokeWithArguments ^MethodHandle @qcb-cancel [handle (long native-id) s])
          (if (= 4 (.get s I32 0)) :cancelled :cannot-cancel)))
      :not-found))

  (queue-status [_]
    (let [js (vals @job-table)]
      {:total-jobs (count js) :backend-load 0.0 :estimated-wait-time 0})))

(defn create-simulator
  "Drop-in for ideal_simulator/create-simulator (ideal_simulator.clj:181-195)."
  ([] (create-simulator {}))
  ([config] (->B200Simulator config (atom {}))))

(comment
  ;; JNA fallback for JDK < 22: same entry points through com.sun.jna.Function
  (import '(com.sun.jna Function NativeLibrary))
  (def lib (NativeLibrary/getInstance "qcb200"))
  (defn jna-call [name & args] (.invokeInt (.getFunction lib name) (to-array args))))
